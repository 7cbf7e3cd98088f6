#!/bin/bash
# ROUND 2, GPU call 3 (1 GPU): the GPU suite with the new default (7 slices of 8 bits), fewer slices for the gradient products, U2 on
# int8, the int8 prediction GEMM, the default bench line, the launch list (n = 20 000: ncu serialises ~2000 launches quickly there).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2d_pytest.log
OZ_TIME_S=0,-1 timeout 300 python scripts/oz_check.py 700 -- 20000 50000 > $O/r2d_oz_default.log 2>&1; echo "default rc=$?"; grep "^time" $O/r2d_oz_default.log
GPSS_OZAKI_GRAD=6 OZ_TIME_S=0,-1 timeout 300 python scripts/oz_check.py 700 -- 20000 50000 > $O/r2d_oz_grad6.log 2>&1; echo "grad6 rc=$?"; grep "^time" $O/r2d_oz_grad6.log
GPSS_OZAKI_GRAD=5 OZ_TIME_S=-1 timeout 300 python scripts/oz_check.py 700 -- 50000 > $O/r2d_oz_grad5.log 2>&1; grep "^time" $O/r2d_oz_grad5.log
GPSS_OZ_U2=1 OZ_TIME_S=-1 timeout 300 python scripts/oz_check.py 700 -- 50000 > $O/r2d_oz_u2.log 2>&1; echo "u2 rc=$?"; grep "^time" $O/r2d_oz_u2.log
timeout 300 python scripts/oz_predict_check.py 2000 20000 50000 > $O/r2d_oz_predict.log 2>&1; echo "predict rc=$?"; cat $O/r2d_oz_predict.log | tail -4
timeout 600 python bench.py > $O/r2d_bench_n1.json 2> $O/r2d_bench_n1.err; echo "bench rc=$?"; cut -c1-300 $O/r2d_bench_n1.json; tail -3 $O/r2d_bench_n1.err
timeout 300 python scripts/fit_n50k.py 50000 3 > $O/r2d_fit_n50k.log 2>&1; tail -2 $O/r2d_fit_n50k.log
python bench.py --n 20000 --steps 1 --warmup 1 --no-cpu-baseline --pred-m 0 > $O/r2d_plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r2d_launches_n20k.csv \
    python bench.py --n 20000 --steps 1 --warmup 1 --no-cpu-baseline --pred-m 0 > $O/r2d_ncu_launches.log 2>&1; echo "launch list rc=$?"
gzip -f $O/r2d_launches_n20k.csv
ls -la $O | tail -12
