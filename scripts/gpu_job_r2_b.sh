#!/bin/bash
# ROUND 2, GPU call 1 (1 GPU): the int8 tensor-core GEMM's three issue variants (GPSS_OZ_VARIANT 0 = round-1 kernel, 1 = merged
# N <= 256 instructions, 2 = merged + 32-byte stages): exactness, throughput, int8 micro-peak, ncu --set full of 0 and 1,
# then the GPU suite with the previously gated bit-exact test and phase times at n = 50 000.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2b_smi.txt 2>&1
for v in 1 2; do
  echo "== variant $v" >> $O/r2b_exact.log
  GPSS_OZ_VARIANT=$v timeout 120 bench_micro/ozaki_gemm exact >> $O/r2b_exact.log 2>&1; echo "exact v$v rc=$?"
  GPSS_OZ_VARIANT=$v timeout 120 bench_micro/ozaki_gemm check >> $O/r2b_exact.log 2>&1; echo "check v$v rc=$?"
  GPSS_OZ_VARIANT=$v timeout 120 bench_micro/ozaki_gemm tri >> $O/r2b_exact.log 2>&1; echo "tri v$v rc=$?"
done
tail -4 $O/r2b_exact.log
timeout 120 bench_micro/int8_peak > $O/r2b_int8_peak.txt 2>&1; echo "peak rc=$?"; cat $O/r2b_int8_peak.txt
for v in 0 1 2; do
  for s in 8 7; do
    echo "== variant $v S $s" >> $O/r2b_variants.txt
    GPSS_OZ_VARIANT=$v timeout 120 bench_micro/ozaki_gemm bench 16384 16384 8192 $s 2>&1 | grep oz_gemm >> $O/r2b_variants.txt
  done
done
GPSS_OZ_VARIANT=1 timeout 120 bench_micro/ozaki_gemm bench 16384 16384 49152 8 2>&1 | grep oz_gemm >> $O/r2b_variants.txt
GPSS_OZ_VARIANT=2 timeout 120 bench_micro/ozaki_gemm bench 16384 16384 49152 8 2>&1 | grep oz_gemm >> $O/r2b_variants.txt
cat $O/r2b_variants.txt
for v in 0 1 2; do
  GPSS_OZ_VARIANT=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel -c 1 -f -o $O/r2b_oz_gemm_v${v}_full \
    bench_micro/ozaki_gemm bench 8192 8192 8192 8 > $O/r2b_ncu_v$v.log 2>&1; echo "ncu v$v rc=$?"
done
GPSS_TEST_ROUND2=1 timeout 600 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2b_pytest.log
for v in 0 1 2; do
  GPSS_OZ_VARIANT=$v OZ_TIME_S=8 timeout 200 python scripts/oz_check.py 700 -- 50000 > $O/r2b_oz_50k_v$v.log 2>&1; tail -1 $O/r2b_oz_50k_v$v.log
done
GPSS_OZ_VARIANT=1 OZ_TIME_S=0,8 timeout 200 python scripts/oz_check.py 3000 -- 20000 > $O/r2b_oz_20k.log 2>&1; tail -2 $O/r2b_oz_20k.log
ls -la $O | tail -20
