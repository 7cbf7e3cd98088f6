#!/bin/bash
# ROUND 2, GPU call 10 (2 GPUs): oz_gemm_kernel capped at 168 registers (the critical-path DMMA kernel becomes co-resident for real) and the
# vector solves beside the inverse -- one-GPU timing at n = 50 000 + the tests that cover them, then the 2-GPU Cholesky variants with trace.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|NCCL version'
OZ_TIME_S=-1 timeout 300 python scripts/oz_check.py 700 -- 20000 50000 2>&1 | grep "^time\|^parity" > $O/r2j_n50k_1gpu.log; cat $O/r2j_n50k_1gpu.log
GPSS_NO_SOLVE_OVERLAP=1 OZ_TIME_S=-1 timeout 300 python scripts/oz_check.py 700 -- 50000 2>&1 | grep "^time" > $O/r2j_n50k_1gpu_serial_solves.log; cat $O/r2j_n50k_1gpu_serial_solves.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "solves_beside or default_pipe or config2 or large_n or int8 or fused_panel or cholesky_failure or caching or bound" > $O/r2j_pytest_1gpu.log 2>&1; echo "pytest 1 GPU rc=$?"; tail -3 $O/r2j_pytest_1gpu.log
GPSS_DIST_PHASES=1 timeout 500 $TR --master-port 29511 scripts/dist_check.py 3000 20000 > $O/r2j_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "$F" $O/r2j_dist_check.log | tail -16
GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29512 scripts/dist_check.py 50000 > $O/r2j_n50k_fused.log 2>&1; echo "n50k fused rc=$?"; grep -v "$F" $O/r2j_n50k_fused.log | tail -9
GPSS_DIST_PANEL=0 GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29513 scripts/dist_check.py 50000 > $O/r2j_n50k_old.log 2>&1; echo "n50k old panel rc=$?"; grep -v "$F" $O/r2j_n50k_old.log | tail -9
GPSS_DIST_U2=int8 GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29514 scripts/dist_check.py 50000 > $O/r2j_n50k_fused_u2int8.log 2>&1; echo "n50k fused, U2 int8 rc=$?"; grep -v "$F" $O/r2j_n50k_fused_u2int8.log | tail -9
timeout 400 $TR --master-port 29515 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2j_bench_n2.json 2> $O/r2j_bench_n2.err; echo "bench2 rc=$?"; cut -c1-200 $O/r2j_bench_n2.json
