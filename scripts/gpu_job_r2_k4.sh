#!/bin/bash
# ROUND 2, GPU call 12 (4 GPUs): sanity of the new multi-GPU defaults beyond two ranks (receive-buffer reuse of the U exchange, the
# interleaved inverse) before the single 8-GPU call, and the 4-GPU bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|NCCL version\|OMP_NUM_THREADS\|^$'
GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 400 $TR --master-port 29511 scripts/dist_check.py 20000 50000 > $O/r2k4_dist_check.log 2>&1; echo "dist_check rc=$?"; grep -v "$F\|dist trace" $O/r2k4_dist_check.log | tail -24; grep "dist trace. rank 0" $O/r2k4_dist_check.log | tail -1
GPSS_TRTRI_INTERLEAVE=0 timeout 300 $TR --master-port 29512 scripts/dist_check.py 50000 > $O/r2k4_n50k_no_interleave.log 2>&1; echo "no interleave rc=$?"; grep "wall" $O/r2k4_n50k_no_interleave.log
timeout 400 $TR --master-port 29515 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2k4_bench_n4.json 2> $O/r2k4_bench_n4.err; echo "bench4 rc=$?"; cut -c1-200 $O/r2k4_bench_n4.json
