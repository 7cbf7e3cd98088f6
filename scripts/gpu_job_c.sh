#!/bin/bash
# 8-GPU job: headline bench (one evaluation over 8 GPUs + sharded prediction leg), partitioned storage at n = 200 000
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/c_bench_n8.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/c_bench_n8.log > gpurun_out/c_bench_n8.json; cat gpurun_out/c_bench_n8.json
grep -v "^{" gpurun_out/c_bench_n8.log | grep -v "^W\|^\*\|OMP_NUM" | tail -5
timeout 900 $TR --master-port 29522 scripts/part_check.py 200000 > gpurun_out/c_part_n200k.log 2>&1; echo "part rc=$?"
grep -v "^W\|^\*\|OMP_NUM" gpurun_out/c_part_n200k.log | tail -12
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3
