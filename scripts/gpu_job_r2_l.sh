#!/bin/bash
# ROUND 2, GPU call 12 (8 GPUs, ONE call: it is charged 8x): BASELINE configs[2] and configs[3] at size, the 8-GPU bench line, parity and the
# phase / critical-path tables of the distributed evaluation at n = 50 000.
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
ITERS=${2:-10}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|NCCL version\|OMP_NUM_THREADS\|^$'
GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29511 scripts/dist_check.py 20000 50000 > $O/r2l_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -v "$F\|dist trace" $O/r2l_dist_check_n$N.log | tail -30
GPSS_DIST_PANEL=fused GPSS_DIST_PHASES=1 timeout 200 $TR --master-port 29512 scripts/dist_check.py 50000 > $O/r2l_dist_check_n${N}_fused.log 2>&1; echo "fused panel rc=$?"; grep "wall\|rank 0 phases\|DIST CHECK" $O/r2l_dist_check_n${N}_fused.log
timeout 400 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $O/r2l_bench_n$N.json 2> $O/r2l_bench_n$N.err; echo "bench rc=$?"; cut -c1-260 $O/r2l_bench_n$N.json
timeout 600 python scripts/fit_n50k.py 50000 $ITERS $N $O/r2l_fit_n50k_${N}gpu.json > $O/r2l_fit_n50k_${N}gpu.log 2>&1; echo "fit rc=$?"; cut -c1-400 $O/r2l_fit_n50k_${N}gpu.log | head -3
timeout 900 bash scripts/predict_10m.sh $N > $O/r2l_predict_10m_${N}gpu.log 2>&1; echo "predict 10M rc=$?"; cat $O/r2l_predict_10m_${N}gpu.log | tail -25
ls -la $O | tail -8
