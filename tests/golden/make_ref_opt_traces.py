"""Generates tests/golden/ref_opt_traces.npz: the probe traces (kind, theta, f, g of every ObjVal / Grad_Values call) of the
UNMODIFIED reference's BFGS and SCG drivers (Opt_pars.cpp:451-538, 979-1124) on the n = 300 synthetic set, recorded by
oracle/_ref/ref_driver --trace.  Needs /root/reference (build container only); the fixture is committed.
   python tests/golden/make_ref_opt_traces.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from gp_ss_ak_b200 import datagen

DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def trace(opt, n, seed, iters):
    X, y = datagen.drillholes(n, seed)
    with tempfile.TemporaryDirectory() as d:
        datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
        out = subprocess.run([DRIVER, "--trace", opt, os.path.join(d, "train.txt"), str(iters), d], capture_output=True, text=True,
                             check=True, env=env, cwd=d).stdout
    kinds, thetas, fs, gs, lines = [], [], [], [], []
    rec = {}
    for line in out.splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "probe":
            kinds.append(0 if tok[2] == "O" else 1)
            thetas.append([float(v) for v in tok[3:13]])
            fs.append(float(tok[14]))
            gs.append([float(v) for v in tok[16:26]] if tok[2] == "G" else [np.nan] * 10)
        elif tok[0] == "Iteration:":
            lines.append(line)
        elif tok[0] in ("theta_start", "theta_fit", "nlml_fit"):
            rec[tok[0]] = np.array(tok[3:], dtype=float)
    rec.update(kind=np.array(kinds), theta=np.array(thetas), f=np.array(fs), g=np.array(gs), iters=iters, stdout_iterations=np.array("\n".join(lines)))
    return rec


if __name__ == "__main__":
    if not os.path.exists(DRIVER):
        raise SystemExit("oracle/_ref/ref_driver is not built (needs /root/reference): make -C oracle")
    out = {}
    for opt, iters in (("BFGS", 12), ("SCG", 14)):
        r = trace(opt, 300, 0, iters)
        for k, v in r.items():
            out["%s_%s" % (opt, k)] = v
        print(opt, "probes", len(r["f"]), "nlml_fit", r["nlml_fit"])
    np.savez_compressed(os.path.join(HERE, "ref_opt_traces.npz"), **out)
