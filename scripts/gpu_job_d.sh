#!/bin/bash
# 8-GPU job: k-chunk sweep of the distributed Cholesky's bulk updates at n = 50k
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for kc in 65536 8192 4096 2048; do
  GPSS_DIST_KCHUNK=$kc GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29531 scripts/dist_check.py 50000 > gpurun_out/d_dist_kc$kc.log 2>&1; echo "kchunk $kc rc=$?"
  grep -E "rep [12]|rank [0-7] phases|CHECK" gpurun_out/d_dist_kc$kc.log | cut -c1-150
done
