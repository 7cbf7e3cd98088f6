#!/bin/bash
# ROUND 2 (2 GPUs): the multi-GPU tests at HEAD (incl. the int8 pipe on a replicated handle and the partitioned variance) and BASELINE configs[0]
# end to end -- the unmodified reference on the host cores against this repo's command line on one GPU, same file, same flags.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q -k "two_gpu or multi or dist or shard" > gpurun_out/r2r_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2r_pytest_multi.log
timeout 300 python scripts/config1_e2e.py 2000 2 > gpurun_out/r2r_config1_e2e.json 2> gpurun_out/r2r_config1_e2e.err; echo "config 1 rc=$?"; cut -c1-1500 gpurun_out/r2r_config1_e2e.json
