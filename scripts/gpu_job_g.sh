#!/bin/bash
# 1-GPU job: full GPU suite with the CUDA-graph path active, small-n timing with / without graphs
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g_pytest.log
python scripts/gpu_perf_check.py 500 2000 4096 8192 > gpurun_out/g_perf_graph.log 2>&1; echo "perf(graph) rc=$?"
GPSS_NO_GRAPH=1 python scripts/gpu_perf_check.py 500 2000 4096 8192 > gpurun_out/g_perf_nograph.log 2>&1; echo "perf(nograph) rc=$?"
echo "--- with graphs"; grep unprofiled gpurun_out/g_perf_graph.log
echo "--- without graphs"; grep unprofiled gpurun_out/g_perf_nograph.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
