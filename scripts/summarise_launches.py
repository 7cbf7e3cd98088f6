"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total and share.
   python scripts/summarise_launches.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
unit = None
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v_ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}[u]
    tot[name][0] += 1
    tot[name][1] += v_ms
total = sum(v[1] for v in tot.values())
print("# %s: %d launches, %.1f ms of kernel time (serialised, cold cache: compare SHARES)" % (sys.argv[1], sum(v[0] for v in tot.values()), total))
print("%-60s %8s %12s %8s" % ("kernel", "launches", "total ms", "share"))
for name, (cnt, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %12.2f %7.2f%%" % (name[:60], cnt, ms, 100 * ms / total))
