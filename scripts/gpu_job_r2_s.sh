#!/bin/bash
# ROUND 2 (1 GPU): two-stream inverse on the int8 path -- parity tests, timing at n = 20 000 / 50 000 against GPSS_INV_TWO_STREAMS=0.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "int8 or default_pipe or config2 or large_n or caching or solves_beside or fused_panel" > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2s_pytest.log
OZ_TIME_S=0,-1 OZ_CHECK_S=7 GPSS_OZAKI_BITS=8 timeout 300 python scripts/oz_check.py 2000 5000 -- 20000 50000 2>&1 | grep "^time\|^parity" > $O/r2s_two_streams.log; cat $O/r2s_two_streams.log
GPSS_INV_TWO_STREAMS=0 OZ_TIME_S=-1 timeout 200 python scripts/oz_check.py 700 -- 20000 50000 2>&1 | grep "^time" > $O/r2s_one_stream.log; cat $O/r2s_one_stream.log
