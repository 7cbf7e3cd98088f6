#!/bin/bash
# 4-GPU job: critical-path breakdown of the distributed Cholesky (GPSS_DIST_TRACE) + the bench line at N = 4
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
GPSS_DIST_TRACE=1 GPSS_DIST_PHASES=1 timeout 300 $TR --master-port 29541 scripts/dist_check.py 50000 > gpurun_out/e_dist_trace_n50k.log 2>&1; echo "trace rc=$?"
grep -E "rep [12]|rank [0-3] phases|CHECK" gpurun_out/e_dist_trace_n50k.log | cut -c1-160
grep "dist trace" gpurun_out/e_dist_trace_n50k.log | tail -4
timeout 400 $TR --master-port 29542 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/e_bench_n4.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/e_bench_n4.log > gpurun_out/e_bench_n4.json; cut -c1-900 gpurun_out/e_bench_n4.json
