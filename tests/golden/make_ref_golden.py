"""Generates tests/golden/ref_*.npz from the UNMODIFIED reference (oracle/_ref/ref_driver, built by oracle/Makefile from
/root/reference against the in-repo Armadillo stand-in + scipy's OpenBLAS).  Only runs in the build container (the
reference sources are not present on the GPU box); the fixtures it writes are committed and travel.

    python tests/golden/make_ref_golden.py

Each fixture holds, with 17 significant digits as dumped by the reference's own classes: raw and standardised data,
params, thetas, nlml, g[10], Alpha, diag(K), diag(D2), two columns of K, predictive mean/variance on a test set that
includes training points, the CLI's back-transformed outputs, and the LBFGS probe trace (kind, theta, f, g).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gp_ss_ak_b200 import datagen          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
THETA0 = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])


def parse_dump(text):
    rec, probes = {}, []
    for line in text.splitlines():
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "probe":
            kind = tok[2]
            vals = tok[3:]
            npar = vals.index("f")                      # 10 for Hyb{ExpAns, Bias}; 4 / 5 for Hyb{Exp | RBF, Bias}
            th = np.array(vals[:npar], dtype=float)
            f = float(vals[npar + 1])
            g = np.array(vals[npar + 3:2 * npar + 3], dtype=float) if kind == "G" else np.full(npar, np.nan)
            probes.append((0.0 if kind == "O" else 1.0, th, f, g))
        else:
            r, c = int(tok[1]), int(tok[2])
            rec[tok[0]] = np.array(tok[3:], dtype=float).reshape((c, r)).T if r * c > 1 else float(tok[3])
    if probes:
        rec["probe_kind"] = np.array([p[0] for p in probes])
        rec["probe_theta"] = np.stack([p[1] for p in probes])
        rec["probe_f"] = np.array([p[2] for p in probes])
        rec["probe_g"] = np.stack([p[3] for p in probes])
    return rec


def make(name, n, seed, thetas, lbfgs_iters, n_test=40, n_coincident=10, cli_iters=0, rock=False, kernel="ExpAns"):
    X, y = datagen.drillholes(n, seed)
    Xt, _ = datagen.drillholes(n_test, seed + 50)
    Xt = np.concatenate([Xt, X[:n_coincident]])
    yt = np.concatenate([datagen.grade_at(Xt[:n_test], seed), y[:n_coincident]])
    if rock:        # the 4-column (rock-type) branch of the ExpAns kernel (Kernel.cpp:872-878, 1169-1173, 1246-1255)
        X = datagen.with_rock_column(X, seed)
        Xt = datagen.with_rock_column(Xt, seed)
    with tempfile.TemporaryDirectory() as d:
        datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
        datagen.write_data_file(os.path.join(d, "test.txt"), Xt, yt)
        with open(os.path.join(d, "thetas.txt"), "w") as f:
            for th in thetas:
                f.write(" ".join("%.17g" % v for v in th) + "\n")
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", GPSS_REF_KERNEL=kernel)   # one BLAS thread: bit-reproducible fixtures
        out = subprocess.run([DRIVER, os.path.join(d, "train.txt"), os.path.join(d, "test.txt"), os.path.join(d, "thetas.txt"),
                              str(lbfgs_iters), d], capture_output=True, text=True, check=True, env=env)
        rec = parse_dump(out.stdout)
        model = os.path.join(d, "ref_model")
        if os.path.exists(model):
            rec["model_file_text"] = np.array(open(model).read())
        rec["statistics_file_text"] = np.array(open(model + "_Statistics.txt").read())
        # the reference's own command line, end to end (README.md:45,49): train for a few iterations, then test
        cli = os.path.join(ROOT, "oracle", "_ref", "gp_ss_ak")
        # Opt_Algs::fail_pre_bfgs is never initialised (Opt_pars.h:218) and the CLI allocates the model with `new`:
        # MALLOC_PERTURB_=255 makes glibc fill fresh allocations with 0x00, i.e. the flag starts false -- the same
        # deterministic starting state as oracle/ref_driver.cpp and the product's host code.
        env = dict(env, MALLOC_PERTURB_="255")
        if cli_iters > 0:
            tr = subprocess.run([cli, "-v", "3", "-pm", "1", "train", "-k", kernel, "-kn", "1", "-o", "LBFGS", "-#", str(cli_iters),
                                 os.path.join(d, "train.txt"), os.path.join(d, "cli_model")], capture_output=True, text=True,
                                stdin=subprocess.DEVNULL, env=env, cwd=d)
            te = subprocess.run([cli, "-v", "3", "-pm", "1", "test", os.path.join(d, "test.txt"), os.path.join(d, "cli_model"),
                                 os.path.join(d, "train.txt")], capture_output=True, text=True, stdin=subprocess.DEVNULL, env=env, cwd=d)
            rec["cli_train_stdout"] = np.array(tr.stdout)
            rec["cli_test_stdout"] = np.array(te.stdout)
            rec["cli_model_text"] = np.array(open(os.path.join(d, "cli_model")).read())
            rec["cli_stats_text"] = np.array(open(os.path.join(d, "cli_model_Statistics.txt")).read())
            rec["cli_predict_text"] = np.array(open(os.path.join(d, "cli_model_predict.txt")).read())
            rec["cli_iters"] = cli_iters
        rec["train_file_text"] = np.array(open(os.path.join(d, "train.txt")).read())
        rec["test_file_text"] = np.array(open(os.path.join(d, "test.txt")).read())
    rec["yt_raw"] = yt
    rec["Xt_raw"] = Xt
    rec["lbfgs_iters"] = lbfgs_iters
    rec["kernel"] = np.array(kernel)
    np.savez_compressed(os.path.join(HERE, name), **rec)
    print(name, "nlml", [rec["nlml_%d" % k] for k in range(int(rec["n_theta"]))], "probes", len(rec.get("probe_f", [])))


if __name__ == "__main__":
    if not os.path.exists(DRIVER):
        raise SystemExit("oracle/_ref/ref_driver is not built (needs /root/reference): make -C oracle")
    rng = np.random.default_rng(42)
    th1 = np.clip(THETA0 * rng.uniform(0.8, 1.25, 10), 1e-4, 6.0)
    th2 = np.clip(THETA0 * rng.uniform(0.8, 1.25, 10), 1e-4, 6.0)
    make("ref_n300.npz", 300, 0, [THETA0, th1, th2], lbfgs_iters=30, cli_iters=6)
    make("ref_n1000.npz", 1000, 1, [THETA0, th1], lbfgs_iters=4)
    make("ref_rock_n300.npz", 300, 2, [THETA0, th1, th2], lbfgs_iters=6, cli_iters=3, rock=True)
    make_iso()


THETA0_EXP = np.array([0.5, 0.9, 0.2, 0.016])            # Hayper_Euc_Exp, Sigma_Exp (Kernel.cpp:585-589), Sigma_Bias, sn2
THETA0_RBF = np.array([0.5, 0.9, 0.5, 0.2, 0.016])       # Hayper_Euc_RBF, inverseWidth_RBF, Sigma_RBF (Kernel.cpp:425-431), Sigma_Bias, sn2


def make_iso():
    """The isotropic members of the kernel family: Hyb{Exp, Bias} and Hyb{RBF, Bias}."""
    rng = np.random.default_rng(7)
    for kernel, th0 in (("Exp", THETA0_EXP), ("RBF", THETA0_RBF)):
        ths = [th0] + [np.clip(th0 * rng.uniform(0.8, 1.25, th0.shape[0]), 1e-4, 6.0) for _ in range(2)]
        make("ref_%s_n300.npz" % kernel.lower(), 300, 3, ths, lbfgs_iters=5, cli_iters=3, kernel=kernel)
