#!/bin/bash
# ROUND 2, GPU call 4 (2 GPUs): the int8 pipe on replicated multi-GPU handles (GPSS_OZAKI_DIST) against one GPU, the multi-GPU
# test files the 1-GPU driver run skips, and the 2-GPU bench lines (int8 and DMMA).
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
GPSS_DIST_PHASES=1 timeout 600 $TR --master-port 29511 scripts/dist_check.py 3000 20000 50000 > $O/r2e_dist_check_int8.log 2>&1; echo "dist_check int8 rc=$?"; grep -v "^W\|^\*\*\*" $O/r2e_dist_check_int8.log | tail -22
GPSS_OZAKI_DIST=0 GPSS_DIST_PHASES=1 timeout 600 $TR --master-port 29512 scripts/dist_check.py 20000 > $O/r2e_dist_check_dmma.log 2>&1; echo "dist_check dmma rc=$?"; grep -v "^W\|^\*\*\*" $O/r2e_dist_check_dmma.log | tail -8
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "two_gpu or multi or dist or shard" > $O/r2e_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 $O/r2e_pytest_multi.log
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2e_bench_n2_int8.json 2> $O/r2e_bench_n2_int8.err; echo "bench2 int8 rc=$?"; cut -c1-250 $O/r2e_bench_n2_int8.json
GPSS_OZAKI_DIST=0 timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --pred-m 0 > $O/r2e_bench_n2_dmma.json 2> $O/r2e_bench_n2_dmma.err; echo "bench2 dmma rc=$?"; cut -c1-250 $O/r2e_bench_n2_dmma.json
ls -la $O | tail -8
