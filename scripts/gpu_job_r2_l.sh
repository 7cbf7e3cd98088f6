#!/bin/bash
# ROUND 2, GPU call (8 GPUs, ONE call: it is charged 8x): BASELINE configs[2] and configs[3] at size, the 8-GPU bench line, parity and the
# phase / critical-path tables of the distributed evaluation at n = 50 000, and config 5's predictive variance at n = 200 000.
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-8}
ITERS=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|NCCL version\|OMP_NUM_THREADS\|^$'
GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 200 $TR --master-port 29511 scripts/dist_check.py 20000 50000 > $O/r2l_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -v "$F\|dist trace\|^ [a-zA-Z]" $O/r2l_dist_check_n$N.log | tail -34; grep "dist trace. rank 0" $O/r2l_dist_check_n$N.log | tail -1 | cut -c1-600
GPSS_TRTRI_INTERLEAVE=0 timeout 120 $TR --master-port 29512 scripts/dist_check.py 50000 > $O/r2l_n50k_no_interleave_n$N.log 2>&1; echo "no interleave rc=$?"; grep "wall\|DIST CHECK" $O/r2l_n50k_no_interleave_n$N.log
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $O/r2l_bench_n$N.json 2> $O/r2l_bench_n$N.err; echo "bench rc=$?"; cut -c1-260 $O/r2l_bench_n$N.json
timeout 300 python scripts/fit_n50k.py 50000 $ITERS $N $O/r2l_fit_n50k_${N}gpu.json > $O/r2l_fit_n50k_${N}gpu.log 2>&1; echo "fit rc=$?"; cut -c1-420 $O/r2l_fit_n50k_${N}gpu.log | head -2
timeout 360 bash scripts/predict_10m.sh $N > $O/r2l_predict_10m_${N}gpu.log 2>&1; echo "predict 10M rc=$?"; tail -22 $O/r2l_predict_10m_${N}gpu.log
GPSS_PART_GRAD=1 timeout 240 $TR --master-port 29514 scripts/part_check.py 200000 > $O/r2l_part_check_n200k.log 2>&1; echo "part_check 200k rc=$?"; grep -v "$F\|   g = \|ref g" $O/r2l_part_check_n200k.log | tail -12 | cut -c1-400
ls -la $O | tail -8
