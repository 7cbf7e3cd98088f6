#!/bin/bash
# BASELINE.json configs[3], end to end through the command line: 10 M block-model centroids (250 x 250 x 160 grid) predicted, mean +
# variance, against an n = 50 000 model -- `gp_ss_ak test` (reader -> standardisation -> factorisation -> sharded gpss_predict ->
# de-standardisation -> sort by y -> <model>_predict.txt writer), one process per GPU (scripts/run_dist_cli.sh).
#   scripts/predict_10m.sh NGPUS [NX NY NZ] [N_TRAIN]        GPSS_OZAKI_PREDICT=1 puts the variance GEMM on the int8 tensor cores
# Prints the phase table of rank 0 (GPSS_TIMING) and preds/s over the prediction phase and over the whole command.
set -u
N=${1:-1}; NX=${2:-250}; NY=${3:-250}; NZ=${4:-160}; NTR=${5:-50000}
HERE="$(cd "$(dirname "$0")/.." && pwd)"
W=$(mktemp -d /tmp/gpss_p10m.XXXXXX)
python - "$W" "$NTR" <<'PY'
import sys, time
sys.path.insert(0, ".")
from gp_ss_ak_b200 import datagen
w, n = sys.argv[1], int(sys.argv[2])
X, y = datagen.drillholes(n, 0)
datagen.write_data_file(w + "/train.txt", X, y)
PY
t0=$(date +%s.%N)
"$HERE/gp_ss_ak_b200/host/tests/make_block_model" "$W/block.txt" $NX $NY $NZ 0 0 0 1000 1000 400
t1=$(date +%s.%N)
echo "block model: $((NX*NY*NZ)) centroids written in $(awk -v a=$t0 -v b=$t1 'BEGIN{printf "%.1f", b-a}') s ($(stat -c %s "$W/block.txt") bytes)"
# a model file: one L-BFGS iteration from the reference's initial parameters (the fit itself is configs[2], scripts/fit_n50k.py)
"$HERE/scripts/run_dist_cli.sh" $N -v 3 -pm 1 train -k ExpAns -kn 1 -o LBFGS -# 1 "$W/train.txt" "$W/model" < /dev/null > "$W/train.log" 2>&1
echo "train rc=$? : $(grep -c . "$W/train.log") lines; model: $(wc -l < "$W/model") lines"
t2=$(date +%s.%N)
GPSS_TIMING=1 "$HERE/scripts/run_dist_cli.sh" $N -v 1 -pm 1 test "$W/block.txt" "$W/model" "$W/train.txt" "$W/predict.txt" < /dev/null > "$W/test.log" 2> "$W/test.err"
rc=$?
t3=$(date +%s.%N)
cat "$W/test.err" | grep "gpss timing"
tail -3 "$W/test.log"
M=$((NX*NY*NZ))
PRED=$(grep "prediction (Calc_Out)" "$W/test.err" | awk '{print $5}')
awk -v m=$M -v n=$N -v a=$t2 -v b=$t3 -v p="$PRED" -v rc=$rc 'BEGIN{printf "test rc=%d : %d points on %d GPU(s): whole command %.1f s = %.0f preds/s wall; prediction phase %.1f s = %.0f preds/s\n", rc, m, n, b-a, m/(b-a), p, (p>0)?m/p:0}'
echo "predict file: $(wc -l < "$W/predict.txt") lines, $(stat -c %s "$W/predict.txt") bytes; first rows:"; head -3 "$W/predict.txt"
rm -rf "$W"
exit $rc
