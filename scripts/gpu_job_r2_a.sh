#!/bin/bash
# ROUND 2, first GPU call (1 GPU, ~6 min): everything the int8 tensor-core path still owes.
#   1. the whole GPU suite with the new default (int8 above n_pad = 8192) + the gated bit-exact test of gpss_test_oz_gemm
#   2. the default bench line (3 + 3 steps) -> profiles/
#   3. the launch list of the same command under ncu, then ncu --set full of oz_gemm_kernel (harness bench, one launch)
#   4. S = 7 vs 8 at n = 50 000 and the opt-in int8 prediction GEMM (GPSS_OZAKI_PREDICT=1) against the DMMA prediction
set -u
mkdir -p gpurun_out
GPSS_TEST_ROUND2=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest.log
timeout 300 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2a_bench_n1.json
GPSS_OZAKI_PREDICT=1 timeout 120 python scripts/oz_predict_check.py > gpurun_out/r2a_oz_predict.log 2>&1; echo "predict rc=$?"; tail -6 gpurun_out/r2a_oz_predict.log
OZ_TIME_S=7,8 timeout 120 python scripts/oz_check.py 700 -- 50000 > gpurun_out/r2a_oz_50k.log 2>&1; tail -3 gpurun_out/r2a_oz_50k.log
GPSS_OZAKI_GRAD=7 OZ_CHECK_S=8 OZ_TIME_S=8 timeout 120 python scripts/oz_check.py 2000 5000 -- 50000 > gpurun_out/r2a_oz_grad7.log 2>&1; tail -5 gpurun_out/r2a_oz_grad7.log
# 8-bit digits (GPSS_OZAKI_BITS=8): 7 slices = the accuracy of 8 x 7 bits with 28 products instead of 36 (CPU study: tests/test_ozaki_cpu.py)
GPSS_OZAKI_BITS=8 OZ_CHECK_S=7,6 OZ_TIME_S=7,6 timeout 150 python scripts/oz_check.py 2000 5000 -- 20000 50000 > gpurun_out/r2a_oz_bits8.log 2>&1; tail -8 gpurun_out/r2a_oz_bits8.log
# BASELINE configs[2] on one GPU: the command line's LBFGS fit at n = 50 000, 3 iterations, end to end
timeout 400 python scripts/fit_n50k.py 50000 3 > gpurun_out/r2a_fit_n50k.log 2>&1; tail -2 gpurun_out/r2a_fit_n50k.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pred-m 0 > gpurun_out/r2a_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2a_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pred-m 0 > gpurun_out/r2a_ncu_launches.log 2>&1
bench_micro/ozaki_gemm bench 16384 16384 8192 8 > gpurun_out/r2a_oz_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -c 1 -o gpurun_out/r2a_oz_gemm_full \
    bench_micro/ozaki_gemm bench 16384 16384 8192 8 > gpurun_out/r2a_ncu_oz.log 2>&1
ls -la gpurun_out | tail -12
# 5. (2 GPUs, separate call: gpurun --gpus 2) the opt-in distributed int8 path against one GPU:
#    GPSS_OZAKI=8 GPSS_OZAKI_DIST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 3000 20000
#    GPSS_OZAKI_DIST=1 python -m torch.distributed.run ... bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline
