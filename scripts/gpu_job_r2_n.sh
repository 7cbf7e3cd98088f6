#!/bin/bash
# ROUND 2 (4 GPUs): do shorter-lived bulk CTAs help the critical path of the distributed Cholesky once the inverse runs inside it?
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 scripts/dist_sweep.py 50000 "" "GPSS_INV_KCHUNK=4096" "GPSS_INV_KCHUNK=4096 GPSS_DIST_KCHUNK=4096" "GPSS_INV_KCHUNK=8192 GPSS_DIST_KCHUNK=8192" "GPSS_INV_KCHUNK=2048 GPSS_DIST_KCHUNK=4096" "GPSS_OZ_U2=0" "GPSS_OZ_U2=0 GPSS_INV_KCHUNK=4096 GPSS_DIST_KCHUNK=4096" "GPSS_TRTRI_INTERLEAVE=0" "" > gpurun_out/r2n_sweep_4gpu.log 2>&1; echo "sweep rc=$?"; grep "^n " gpurun_out/r2n_sweep_4gpu.log
