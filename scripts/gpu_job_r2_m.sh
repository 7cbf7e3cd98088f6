#!/bin/bash
# ROUND 2, GPU call 13 (1 GPU): BASELINE configs[2] on one GPU (same iteration count as the 8-GPU run, for the trajectory comparison), the DRAM
# bytes of the B^-1 = U U^T launch at n = 50 000 with the 7 x 8-bit default, the whole -m gpu suite and the default bench line at HEAD.
set -u
mkdir -p gpurun_out
O=gpurun_out
ITERS=${1:-10}
timeout 900 python scripts/fit_n50k.py 50000 $ITERS 1 $O/r2m_fit_n50k_1gpu.json > $O/r2m_fit_n50k_1gpu.log 2>&1; echo "fit rc=$?"; cut -c1-400 $O/r2m_fit_n50k_1gpu.log | head -3
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2m_pytest.log
timeout 900 python bench.py > $O/r2m_bench_n1.json 2> $O/r2m_bench_n1.err; echo "bench rc=$?"; cut -c1-300 $O/r2m_bench_n1.json
timeout 300 python bench.py --impl reference > $O/r2m_bench_reference.json 2> $O/r2m_bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 $O/r2m_bench_reference.json
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:oz_gemm_kernel --clock-control none --csv \
    --log-file $O/r2m_oz_dram.csv python scripts/profile_target.py 50000 256 > $O/r2m_ncu_dram.log 2>&1; echo "dram pass rc=$?"
gzip -f $O/r2m_oz_dram.csv
ls -la $O | tail -8
