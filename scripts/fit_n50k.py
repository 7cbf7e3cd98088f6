"""BASELINE.json configs[2] on one GPU: the reference's own command line (`gp_ss_ak -v 3 -pm 1 train -k ExpAns -kn 1 -o LBFGS -# ITERS`)
through the host classes on a synthetic n-point drillhole file, timed end to end (file read, standardisation, every optimiser probe).

    python scripts/fit_n50k.py [n=50000] [iters=3]
"""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with tempfile.TemporaryDirectory() as d:
    X, y = datagen.drillholes(n, 0)
    t0 = time.perf_counter()
    datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
    t_write = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = subprocess.run([CLI, "-v", "3", "-pm", "1", "train", "-k", "ExpAns", "-kn", "1", "-o", "LBFGS", "-#", str(iters),
                          os.path.join(d, "train.txt"), os.path.join(d, "model")], capture_output=True, text=True, stdin=subprocess.DEVNULL, cwd=d)
    wall = time.perf_counter() - t0
    ll = re.findall(r"Log likelihood:\s*(-?[0-9.eE+-]+)", out.stdout)
    its = re.findall(r"-logL:\s*(-?[0-9.eE+-]+)", out.stdout)
    print("fit n %d, %d LBFGS iterations: rc %d, wall %.1f s (data file written in %.1f s); -logL %s -> %s; per-iteration %s"
          % (n, iters, out.returncode, wall, t_write, ll[0] if ll else "?", ll[-1] if ll else "?", its), flush=True)
    if out.returncode != 0:
        print(out.stdout[-2000:], out.stderr[-2000:])
    sys.exit(out.returncode)
