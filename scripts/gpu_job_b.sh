#!/bin/bash
# 2-GPU job: multi-GPU parity tests, distributed evaluation at n = 50k with phase times, partitioned storage at n = 20k
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu or split_k" > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
GPSS_DIST_PHASES=1 timeout 600 $TR --master-port 29511 scripts/dist_check.py 50000 > gpurun_out/b_dist_n50k.log 2>&1; echo "dist rc=$?"
grep -v "^W\|^\*" gpurun_out/b_dist_n50k.log | tail -12
timeout 600 $TR --master-port 29512 scripts/part_check.py 20000 > gpurun_out/b_part_n20k.log 2>&1; echo "part rc=$?"
grep -v "^W\|^\*" gpurun_out/b_part_n20k.log | tail -12
