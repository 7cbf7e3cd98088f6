#!/bin/bash
# ROUND 2, GPU call 8 (1 GPU, after the container was re-created): the whole -m gpu suite at HEAD (incl. the two-member sums), smoke(),
# the default bench line, the launch list of the bench command, and ncu --set full of oz_gemm_kernel (library at n = 20 000 + the harness).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2h_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/r2h_smoke.log 2>&1; echo "smoke rc=$?"; grep "^smoke" $O/r2h_smoke.log
timeout 900 python bench.py > $O/r2h_bench_n1.json 2> $O/r2h_bench_n1.err; echo "bench rc=$?"; cut -c1-400 $O/r2h_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/r2h_launches_bench_n50k.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --pred-m 0 > $O/r2h_ncu_launches.log 2>&1; echo "launch list rc=$?"
gzip -f $O/r2h_launches_bench_n50k.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel --launch-skip 106 --launch-count 12 \
    -f -o $O/r2h_oz_gemm_lib_n20k python scripts/profile_target.py 20000 256 > $O/r2h_ncu_full_lib.log 2>&1; echo "ncu full (library) rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:oz_gemm_kernel --launch-skip 2 --launch-count 1 \
    -f -o $O/r2h_oz_gemm_harness bench_micro/ozaki_gemm bench 16384 16384 16384 7 > $O/r2h_ncu_full_harness.log 2>&1; echo "ncu full (harness) rc=$?"
ls -la $O | tail -12
