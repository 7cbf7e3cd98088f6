#!/bin/bash
# ROUND 2, GPU call 6 (1 GPU): k-segments inside one launch (8-bit digits) -- bit-exact tests, parity, timing at n = 50 000 with the
# whole-stage (1) and per-plane-unit (3) barrier variants.
set -u
mkdir -p gpurun_out
O=gpurun_out
for v in 1 3; do
  GPSS_OZ_VARIANT=$v timeout 120 bench_micro/ozaki_gemm exact > $O/r2g_exact_v$v.log 2>&1; echo "exact v$v rc=$?"
  GPSS_OZ_VARIANT=$v timeout 400 python -m pytest tests -m gpu -q -k "int8 or bit_exact or default_pipe or large_n" > $O/r2g_pytest_v$v.log 2>&1; echo "pytest v$v rc=$?"; tail -2 $O/r2g_pytest_v$v.log
  GPSS_OZ_VARIANT=$v OZ_TIME_S=0,-1 timeout 300 python scripts/oz_check.py 2000 -- 20000 50000 2>&1 | grep "^time\|^parity" > $O/r2g_n50k_v$v.log; cat $O/r2g_n50k_v$v.log
done
ls -la $O | tail -6
