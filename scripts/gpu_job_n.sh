#!/bin/bash
# last GPU seconds of round 1: the default bench line with the int8 path (short form), then the new int8 parity tests
set -u
mkdir -p gpurun_out
timeout -s KILL 36 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pred-m 8192 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/n_bench.json; tail -3 gpurun_out/n_bench.err
timeout -s KILL 30 python -m pytest tests/test_gpu_parity.py -x -q -k "int8_tensor" > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/n_pytest.log
