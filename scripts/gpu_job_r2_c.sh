#!/bin/bash
# ROUND 2, GPU call 2 (1 GPU): the GPU suite with the new parity tests of the default int8 pipe; 8-bit digits (7 slices, 28 products)
# against the DMMA pipe and the bit-exact restatement; S = 8 vs DMMA at n = 50 000 (nlml, g, alpha); the default bench line; launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2c_pytest.log
GPSS_OZAKI_BITS=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact" > $O/r2c_pytest_bits8.log 2>&1; echo "pytest bits8 rc=$?"; tail -2 $O/r2c_pytest_bits8.log
timeout 120 bench_micro/ozaki_gemm bench 16384 16384 8192 8 2>&1 | grep oz_gemm > $O/r2c_micro.txt
timeout 120 bench_micro/ozaki_gemm bench 16384 16384 49152 8 2>&1 | grep oz_gemm >> $O/r2c_micro.txt; cat $O/r2c_micro.txt
GPSS_OZAKI_BITS=8 OZ_CHECK_S=7 OZ_TIME_S=0,7 timeout 400 python scripts/oz_check.py 2000 5000 -- 20000 50000 > $O/r2c_oz_bits8.log 2>&1; echo "bits8 rc=$?"; tail -8 $O/r2c_oz_bits8.log
OZ_CHECK_S=8 OZ_TIME_S=0,8,7 timeout 400 python scripts/oz_check.py 700 -- 50000 > $O/r2c_oz_50k.log 2>&1; echo "s8 rc=$?"; tail -4 $O/r2c_oz_50k.log
timeout 600 python bench.py > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $O/r2c_bench_n1.json; tail -3 $O/r2c_bench_n1.err
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pred-m 0 > $O/r2c_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/r2c_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pred-m 0 > $O/r2c_ncu_launches.log 2>&1; echo "launch list rc=$?"
gzip -f $O/r2c_launches.csv
# DRAM bytes of every oz_gemm_kernel launch of one evaluation at n = 50 000 (single-pass metrics: no replay, no memory save/restore)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:oz_gemm_kernel --clock-control none --csv \
    --log-file $O/r2c_oz_dram.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --pred-m 0 > $O/r2c_ncu_dram.log 2>&1; echo "dram pass rc=$?"
gzip -f $O/r2c_oz_dram.csv
timeout 60 bench_micro/int8_peak > $O/r2c_int8_peak.txt 2>&1
ls -la $O | tail -12
