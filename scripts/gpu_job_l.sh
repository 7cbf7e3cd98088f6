#!/bin/bash
# last GPU seconds of round 1: the two new configs[0] tests, then the experimental int8 (Ozaki) GEMM harness, each bounded
set -u
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_parity.py::test_against_compiled_reference tests/test_gpu_cli.py::test_cli_baseline_config0_n2000 -x -q > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
timeout -s KILL 15 bench_micro/ozaki_gemm exact > gpurun_out/l_oz_exact.log 2>&1; rc=$?; echo "oz exact rc=$rc"; tail -12 gpurun_out/l_oz_exact.log
if [ $rc -eq 0 ]; then
  timeout -s KILL 20 bench_micro/ozaki_gemm bench 16384 16384 8192 7 > gpurun_out/l_oz_bench7.log 2>&1; echo "oz bench rc=$?"; tail -4 gpurun_out/l_oz_bench7.log
  timeout -s KILL 20 bench_micro/ozaki_gemm check > gpurun_out/l_oz_check.log 2>&1; echo "oz check rc=$?"; tail -4 gpurun_out/l_oz_check.log
  timeout -s KILL 20 bench_micro/ozaki_gemm bench 16384 16384 8192 8 > gpurun_out/l_oz_bench8.log 2>&1; tail -2 gpurun_out/l_oz_bench8.log
fi
nvidia-smi --query-gpu=name,clocks.sm,memory.used --format=csv,noheader
