#!/bin/bash
# opt-in int8 path (GPSS_OZAKI) integrated in libgpss.so: harness regression, parity against the default path, phase times
set -u
mkdir -p gpurun_out
timeout -s KILL 15 bench_micro/ozaki_gemm exact > gpurun_out/m_oz_exact.log 2>&1; echo "oz exact rc=$?"; grep -c "0 mismatches" gpurun_out/m_oz_exact.log
timeout -s KILL 15 bench_micro/ozaki_gemm tri > gpurun_out/m_oz_tri.log 2>&1; echo "oz tri rc=$?"; cat gpurun_out/m_oz_tri.log
timeout -s KILL 60 python scripts/oz_check.py 2000 5000 -- 20000 > gpurun_out/m_oz_check.log 2>&1; echo "oz_check rc=$?"; tail -12 gpurun_out/m_oz_check.log
OZ_TIME_S=7 timeout -s KILL 60 python scripts/oz_check.py 700 -- 50000 > gpurun_out/m_oz_50k_s7.log 2>&1; echo "50k S7 rc=$?"; tail -3 gpurun_out/m_oz_50k_s7.log
OZ_TIME_S=0 OZ_CHECK_S=8 timeout -s KILL 60 python scripts/oz_check.py 700 -- 50000 > gpurun_out/m_oz_50k_s0.log 2>&1; echo "50k S0 rc=$?"; tail -2 gpurun_out/m_oz_50k_s0.log
nvidia-smi --query-gpu=name,clocks.sm,memory.used --format=csv,noheader
