"""BASELINE.json configs[2]: the reference's own command line (`gp_ss_ak -v 3 -pm 1 train -k ExpAns -kn 1 -o LBFGS -# ITERS`,
/root/reference/gp_ss_ak.cpp:288-306 driving Opt_pars.cpp:179-332) through the host classes on a synthetic n-point drillhole file,
timed end to end (file read, standardisation, every optimiser probe, model file, fitted values), on 1 GPU or one process per GPU.

    python scripts/fit_n50k.py [n=50000] [iters=3] [gpus=1] [json_out]

Prints one JSON line: wall time, optimiser calls to the device and their device time (GPSS_TIMING), the -logL trajectory and the
fitted parameters read back from the model file -- two runs (say 1 and 8 GPUs) are compared with scripts/fit_compare.py.
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gp_ss_ak_b200 import datagen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "gp_ss_ak_b200", "host", "gp_ss_ak")
LAUNCH = os.path.join(ROOT, "scripts", "run_dist_cli.sh")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
json_out = sys.argv[4] if len(sys.argv) > 4 else None
with tempfile.TemporaryDirectory() as d:
    X, y = datagen.drillholes(n, 0)
    t0 = time.perf_counter()
    datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
    t_write = time.perf_counter() - t0
    args = ["-v", "3", "-pm", "1", "train", "-k", "ExpAns", "-kn", "1", "-o", "LBFGS", "-#", str(iters),
            os.path.join(d, "train.txt"), os.path.join(d, "model")]
    cmd = [CLI] + args if gpus == 1 else [LAUNCH, str(gpus)] + args
    env = dict(os.environ, GPSS_TIMING="1")
    t0 = time.perf_counter()
    out = subprocess.run(cmd, capture_output=True, text=True, stdin=subprocess.DEVNULL, cwd=d, env=env)
    wall = time.perf_counter() - t0
    ll = re.findall(r"Log likelihood:\s*(-?[0-9.eE+-]+)", out.stdout)
    its = re.findall(r"-logL:\s*(-?[0-9.eE+-]+)", out.stdout)
    mt = re.search(r"optimiser: (\d+) objective \+ (\d+) objective-and-gradient calls to the device, ([0-9.]+) s", out.stderr)
    model = {}
    try:
        with open(os.path.join(d, "model")) as f:
            for line in f:
                if "=" in line and not line.startswith("#"):
                    k, v = line.strip().split("=", 1)
                    model.setdefault(k, []).append(v)
    except OSError:
        pass
    res = {"config": "LBFGS fit, ExpAns+Bias, n=%d, -# %d" % (n, iters), "n": n, "iters": iters, "gpus": gpus, "rc": out.returncode,
           "wall_s": round(wall, 2), "data_file_write_s": round(t_write, 2),
           "objective_calls": int(mt.group(1)) if mt else None, "gradient_calls": int(mt.group(2)) if mt else None,
           "device_s": float(mt.group(3)) if mt else None,
           "nlml_first": float(ll[0]) if ll else None, "nlml_last": float(ll[-1]) if ll else None,
           "nlml_per_iteration": [float(v) for v in its], "model": model,
           "timing": [l for l in out.stderr.splitlines() if l.startswith("[gpss timing]")]}
    line = json.dumps(res)
    print(line, flush=True)
    if json_out:
        with open(json_out, "w") as f:
            f.write(line + "\n")
    if out.returncode != 0:
        print(out.stdout[-2000:], out.stderr[-2000:])
    sys.exit(out.returncode)
