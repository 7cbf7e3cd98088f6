#!/bin/bash
# 1-GPU job: ncu --set full of the dominant kernel (standalone representative shape + in-situ tail of an evaluation)
set -u
mkdir -p gpurun_out
bench_micro/gemm_bench 16384 16384 8192 2 > gpurun_out/f_gemm_bench_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_ws -s 1 -c 2 -o gpurun_out/f_gemm_ws_standalone \
    bench_micro/gemm_bench 16384 16384 8192 2 > gpurun_out/f_ncu_gemm_standalone.log 2>&1
echo "ncu standalone rc=$?"; cat gpurun_out/f_gemm_bench_plain.log
python scripts/profile_target.py 20000 8192 > gpurun_out/f_plain_target.log 2>&1 &&
ncu --set full --clock-control none -k regex:gemm_nt_ws -s 930 -c 40 -o gpurun_out/f_gemm_ws_insitu \
    python scripts/profile_target.py 20000 8192 > gpurun_out/f_ncu_gemm_insitu.log 2>&1
echo "ncu insitu rc=$?"; tail -3 gpurun_out/f_ncu_gemm_insitu.log
ls -la gpurun_out | grep " f_"
