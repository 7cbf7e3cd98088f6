#!/bin/bash
# Multi-GPU launch of the host command line: one gp_ss_ak process per GPU (gp_ss_ak_b200/host/DistHost.h).
#   scripts/run_dist_cli.sh N [gp_ss_ak arguments ...]        stdin (the CLI's prompts) is replicated to every rank
# Rank 0 prints and writes the files; the exit status is the first non-zero status of any rank.
set -u
N=$1; shift
HERE="$(cd "$(dirname "$0")/.." && pwd)"
CLI="$HERE/gp_ss_ak_b200/host/gp_ss_ak"
IN=$(mktemp); cat > "$IN" < /dev/stdin 2>/dev/null || true
export GPSS_WORLD=$N GPSS_JOB=${GPSS_JOB:-$$}
pids=()
for r in $(seq 0 $((N - 1))); do
  GPSS_RANK=$r GPSS_DEVICE=$r "$CLI" "$@" < "$IN" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || { s=$?; [ $rc -eq 0 ] && rc=$s; }; done
rm -f "$IN"
exit $rc
