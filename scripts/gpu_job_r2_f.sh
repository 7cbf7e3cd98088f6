#!/bin/bash
# ROUND 2, GPU call 5 (1 GPU): per-plane-unit barriers (GPSS_OZ_VARIANT=3) -- exactness, micro-bench, n = 50 000; DMMA co-residency on/off;
# the tests added since the last run; launch list of one evaluation at n = 20 000; DRAM bytes per oz_gemm launch with the 7 x 8-bit default.
set -u
mkdir -p gpurun_out
O=gpurun_out
for m in exact check tri; do GPSS_OZ_VARIANT=3 timeout 120 bench_micro/ozaki_gemm $m >> $O/r2f_exact_v3.log 2>&1; echo "$m v3 rc=$?"; done
tail -3 $O/r2f_exact_v3.log
for v in 1 3; do for s in 7 8; do for k in 8192 49152; do
  echo "== variant $v S $s" >> $O/r2f_variants.txt
  GPSS_OZ_VARIANT=$v timeout 120 bench_micro/ozaki_gemm bench 16384 16384 $k $s 2>&1 | grep oz_gemm >> $O/r2f_variants.txt
done; done; done
cat $O/r2f_variants.txt
GPSS_OZ_VARIANT=1 OZ_TIME_S=-1 timeout 200 python scripts/oz_check.py 700 -- 50000 2>&1 | grep "^time" > $O/r2f_n50k.log
GPSS_OZ_VARIANT=3 OZ_TIME_S=-1 timeout 200 python scripts/oz_check.py 700 -- 50000 2>&1 | grep "^time" >> $O/r2f_n50k.log
GPSS_OZ_VARIANT=3 GPSS_DMMA_CORESIDENT=0 OZ_TIME_S=-1 timeout 200 python scripts/oz_check.py 700 -- 50000 2>&1 | grep "^time" >> $O/r2f_n50k.log
GPSS_OZ_VARIANT=3 OZ_TIME_S=0,-1 timeout 200 python scripts/oz_check.py 2000 5000 -- 20000 > $O/r2f_v3_parity.log 2>&1; echo "v3 parity rc=$?"
cat $O/r2f_n50k.log; tail -3 $O/r2f_v3_parity.log
timeout 600 python -m pytest tests -m gpu -q -k "gemm or int8 or white or rejects" > $O/r2f_pytest_subset.log 2>&1; echo "pytest subset rc=$?"; tail -3 $O/r2f_pytest_subset.log
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $O/r2f_launches_n20k.csv \
    python scripts/profile_target.py 20000 2048 > $O/r2f_ncu_launches.log 2>&1; echo "launch list rc=$?"
gzip -f $O/r2f_launches_n20k.csv
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:oz_gemm_kernel --clock-control none --csv \
    --log-file $O/r2f_oz_dram.csv python scripts/profile_target.py 50000 256 > $O/r2f_ncu_dram.log 2>&1; echo "dram pass rc=$?"
gzip -f $O/r2f_oz_dram.csv
ls -la $O | tail -10
