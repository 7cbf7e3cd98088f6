#!/bin/bash
# 1-GPU job: GPU test suite, default bench, ncu launch list of the bench command, ncu --set full of the path's kernels
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/a_pytest_gpu.log
tail -3 gpurun_out/a_pytest_gpu.log
python bench.py > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; echo "bench rc=$?"
cat gpurun_out/a_bench_n1.json
python bench.py --impl reference > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/a_bench_ref.json
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pred-m 8192 > gpurun_out/a_plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pred-m 8192 > gpurun_out/a_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/profile_target.py 20000 8192 > gpurun_out/a_plain_target.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_ws -s 1090 -c 40 -o gpurun_out/a_gemm_ws_insitu \
    python scripts/profile_target.py 20000 8192 > gpurun_out/a_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'kbuild_lower|grad_pass|cross_build|var_finish|kmatvec|mean_finish' -c 10 \
    -o gpurun_out/a_other_kernels python scripts/profile_target.py 20000 8192 > gpurun_out/a_ncu_other.log 2>&1
echo "ncu other rc=$?"
ls -la gpurun_out | tail -20
