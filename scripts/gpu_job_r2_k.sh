#!/bin/bash
# ROUND 2, GPU call 11 (2 GPUs): the inverse issued inside the distributed Cholesky (st8 / st9) and the trapezoid, pipelined exchange of
# the U slices -- parity against one GPU, A/B against GPSS_TRTRI_INTERLEAVE=0, the multi-GPU tests, the 2-GPU bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|NCCL version\|OMP_NUM_THREADS\|^$'
GPSS_DIST_PHASES=1 timeout 600 $TR --master-port 29511 scripts/dist_check.py 3000 20000 50000 > $O/r2k_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "$F" $O/r2k_dist_check.log | tail -26
GPSS_TRTRI_INTERLEAVE=0 timeout 300 $TR --master-port 29512 scripts/dist_check.py 50000 > $O/r2k_n50k_no_interleave.log 2>&1; echo "no interleave rc=$?"; grep -v "$F" $O/r2k_n50k_no_interleave.log | tail -5
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "two_gpu or multi or dist or shard" > $O/r2k_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/r2k_pytest_multi.log
timeout 400 $TR --master-port 29515 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2k_bench_n2.json 2> $O/r2k_bench_n2.err; echo "bench2 rc=$?"; cut -c1-200 $O/r2k_bench_n2.json
