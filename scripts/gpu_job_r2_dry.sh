#!/bin/bash
# dry run (1 GPU, small sizes) of the two command-line scripts the 8-GPU call will run at size
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python scripts/fit_n50k.py 6000 2 1 $O/dry_fit.json > $O/dry_fit.log 2>&1; echo "fit rc=$?"; cut -c1-600 $O/dry_fit.log | head -4
timeout 300 bash scripts/predict_10m.sh 1 100 100 50 20000 > $O/dry_predict.log 2>&1; echo "predict rc=$?"; tail -22 $O/dry_predict.log
