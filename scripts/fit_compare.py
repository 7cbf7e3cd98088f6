"""Two runs of scripts/fit_n50k.py (say 1 GPU and 8 GPUs) side by side: probes, wall time, and whether the optimiser walked the same
trajectory -- the -logL printed after every iteration and the parameters of the model file (6 significant digits, like the reference's
writer, GP_Utils.cpp:1324-1425).     python scripts/fit_compare.py a.json b.json"""
import json
import sys

a, b = (json.loads(open(p).readline()) for p in sys.argv[1:3])
print("%-28s %14s %14s" % ("", "%d GPU(s)" % a["gpus"], "%d GPU(s)" % b["gpus"]))
for k in ("wall_s", "device_s", "objective_calls", "gradient_calls", "nlml_first", "nlml_last"):
    print("%-28s %14s %14s" % (k, a[k], b[k]))
ta, tb = a["nlml_per_iteration"], b["nlml_per_iteration"]
m = min(len(ta), len(tb))
worst = max((abs(x - y) / max(1.0, abs(x)) for x, y in zip(ta[:m], tb[:m])), default=0.0)
print("iterations printed: %d / %d; largest relative difference of -logL along the trajectory: %.2e" % (len(ta), len(tb), worst))
same = True
for k in sorted(set(a["model"]) | set(b["model"])):
    if a["model"].get(k) != b["model"].get(k):
        same = False
        print("model file differs at %s: %s | %s" % (k, a["model"].get(k), b["model"].get(k)))
print("model files (parameters as written, 6 significant digits):", "IDENTICAL" if same else "DIFFERENT")
sys.exit(0 if (worst < 1e-6 and a["gradient_calls"] == b["gradient_calls"]) else 1)
