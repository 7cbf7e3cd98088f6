#!/bin/bash
# ROUND 2: wave-aware row partition of the int8 inverse (kind 2).  N = 2: parity against one GPU.  N = 8: A/B against the flop-balanced
# partition and against the interleaving, one handle per partition (scripts/dist_sweep.py).
set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|NCCL version\|OMP_NUM_THREADS\|^$'
if [ "$N" = "2" ]; then
  GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29541 scripts/dist_check.py 3000 20000 > gpurun_out/r2q_dist_check_n2.log 2>&1; echo "dist_check rc=$?"; grep -v "$F" gpurun_out/r2q_dist_check_n2.log | tail -14
else
  timeout 200 $TR --master-port 29542 scripts/dist_sweep.py 50000 "" "GPSS_TRTRI_INTERLEAVE=0" "GPSS_TRTRI_INTERLEAVE=1" > gpurun_out/r2q_sweep_n${N}_kind2.log 2>&1; echo "kind 2 rc=$?"; grep "^n " gpurun_out/r2q_sweep_n${N}_kind2.log
  GPSS_UROW_KIND=0 timeout 200 $TR --master-port 29543 scripts/dist_sweep.py 50000 "" "GPSS_TRTRI_INTERLEAVE=0" > gpurun_out/r2q_sweep_n${N}_kind0.log 2>&1; echo "kind 0 rc=$?"; grep "^n " gpurun_out/r2q_sweep_n${N}_kind0.log
fi
