#!/bin/bash
# ROUND 2, last 1-GPU call: the -m gpu suite and smoke() at HEAD
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; grep "^smoke" gpurun_out/r2t_smoke.log
