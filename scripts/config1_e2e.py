"""BASELINE.json configs[0] end to end on ONE box: `gp_ss_ak -v 3 -pm 1 train -k ExpAns -kn 1 -o LBFGS -# ITERS` on the same synthetic
n = 2000 drillhole file, run once by the UNMODIFIED reference compiled into oracle/_ref (the box's host cores) and once by this repo's
command line (one B200), wall time of each whole command and the printed results side by side.  The only configuration the reference can
run at its own size, hence the only speed-up that is measured rather than extrapolated.

    python scripts/config1_e2e.py [n=2000] [iters=2]
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gp_ss_ak_b200 import datagen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
REF = os.path.join(ROOT, "oracle", "_ref", "gp_ss_ak")
OURS = os.path.join(ROOT, "gp_ss_ak_b200", "host", "gp_ss_ak")
res = {"config": "gp_ss_ak -v 3 -pm 1 train -k ExpAns -kn 1 -o LBFGS -# %d, n = %d" % (iters, n), "host_cores": os.cpu_count()}
with tempfile.TemporaryDirectory() as d:
    X, y = datagen.drillholes(n, 0)
    datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
    for name, exe in (("reference_cpu", REF), ("this_repo_gpu", OURS), ("this_repo_gpu_second_run", OURS)):
        if not os.path.exists(exe):
            res[name] = {"unavailable": exe}
            continue
        t0 = time.perf_counter()
        out = subprocess.run([exe, "-v", "3", "-pm", "1", "train", "-k", "ExpAns", "-kn", "1", "-o", "LBFGS", "-#", str(iters),
                              os.path.join(d, "train.txt"), os.path.join(d, "model_" + name)], capture_output=True, text=True,
                             stdin=subprocess.DEVNULL, cwd=d, env=dict(os.environ, GPSS_TIMING="1"))
        wall = time.perf_counter() - t0
        ll = re.findall(r"Log likelihood:\s*(-?[0-9.eE+-]+)", out.stdout)
        mse = re.findall(r"Mean Square Error of training:\s*([0-9.eE+-]+)", out.stdout)
        pars = re.findall(r"^(\w+_ExpAns|Sigma_Bias): (\S+)$", out.stdout, re.M)
        calls = re.search(r"optimiser: (\d+) objective \+ (\d+) objective-and-gradient calls to the device, ([0-9.]+) s", out.stderr)
        res[name] = {"rc": out.returncode, "wall_s": round(wall, 3), "log_likelihood_lines": ll, "mse_train": mse[-1] if mse else None,
                     "final_parameters": pars[-9:], "device_calls": [int(calls.group(1)), int(calls.group(2))] if calls else None,
                     "device_s": float(calls.group(3)) if calls else None}
if "wall_s" in res.get("reference_cpu", {}) and "wall_s" in res.get("this_repo_gpu_second_run", {}):
    res["speedup_whole_command"] = round(res["reference_cpu"]["wall_s"] / res["this_repo_gpu_second_run"]["wall_s"], 1)
    res["same_printed_results"] = (res["reference_cpu"]["log_likelihood_lines"] == res["this_repo_gpu"]["log_likelihood_lines"]
                                   and res["reference_cpu"]["final_parameters"] == res["this_repo_gpu"]["final_parameters"])
print(json.dumps(res), flush=True)
