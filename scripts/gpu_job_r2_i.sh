#!/bin/bash
# ROUND 2, GPU call 9 (2 GPUs): the fused distributed panel (diagonal-block chain + one full-height GEMM, U2 beside the chain) against the
# 16-launch panel, with GPSS_DIST_TRACE; parity against one GPU; the multi-GPU tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|NCCL version'
GPSS_DIST_PHASES=1 timeout 500 $TR --master-port 29511 scripts/dist_check.py 3000 20000 > $O/r2i_dist_check_fused.log 2>&1; echo "dist_check fused rc=$?"; grep -v "$F" $O/r2i_dist_check_fused.log | tail -16
GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29512 scripts/dist_check.py 50000 > $O/r2i_n50k_fused.log 2>&1; echo "n50k fused rc=$?"; grep -v "$F" $O/r2i_n50k_fused.log | tail -12
GPSS_DIST_PANEL=0 GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29513 scripts/dist_check.py 50000 > $O/r2i_n50k_old.log 2>&1; echo "n50k old panel rc=$?"; grep -v "$F" $O/r2i_n50k_old.log | tail -12
GPSS_DIST_U2=int8 GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29514 scripts/dist_check.py 50000 > $O/r2i_n50k_fused_u2int8.log 2>&1; echo "n50k fused, U2 int8 rc=$?"; grep -v "$F" $O/r2i_n50k_fused_u2int8.log | tail -12
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "two_gpu or multi or dist or shard" > $O/r2i_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/r2i_pytest_multi.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_panel" > $O/r2i_pytest_fused1.log 2>&1; echo "pytest fused panel on one GPU rc=$?"; tail -3 $O/r2i_pytest_fused1.log
