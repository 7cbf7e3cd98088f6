#!/bin/bash
# final 1-GPU verification: full GPU suite, smoke, default bench line
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/k_bench_n1.json 2> gpurun_out/k_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/k_bench_n1.json
