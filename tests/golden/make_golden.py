"""Generates tests/golden/*.npz with the CPU oracle (oracle/gpss_oracle.py).

The reference ships no golden vectors and cannot be built here (Armadillo absent), so these fixtures pin the
ORACLE's output (reference-literal path: IRLS + Brent + 3 Cholesky + matrix-form gradients) at the time the
restatement was reviewed line by line against the reference; they guard against regressions of the oracle and
serve as known answers for the CUDA path.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gp_ss_ak_b200 import datagen          # noqa: E402
from oracle import gpss_oracle as O        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(n, seed, name, thetas):
    X, y = datagen.drillholes(n, seed)
    Xs, ys, params, st = O.standardise_train(X, y)
    Xt_raw, _ = datagen.drillholes(40, seed + 50)
    Xt_raw = np.concatenate([Xt_raw, X[:10]], axis=0)          # 10 test points coincide with training points
    Xt, _ = O.apply_standardise(Xt_raw, np.zeros(Xt_raw.shape[0]), params)
    rec = {"X_raw": X, "y_raw": y, "Xs": Xs, "ys": ys, "params": params, "Xt": Xt, "thetas": np.array(thetas)}
    for k, th in enumerate(thetas):
        L, g, gp = O.nlml_and_grad(Xs, ys, th, dist="defined", literal=True)
        mu, var = gp.predict(Xt)
        rng = np.random.default_rng(k)
        ii = rng.integers(0, n, 200)
        jj = rng.integers(0, n, 200)
        rec["nlml_%d" % k] = L
        rec["g_%d" % k] = g
        rec["alpha_%d" % k] = gp.Alpha
        rec["mu_%d" % k] = mu
        rec["var_%d" % k] = var
        rec["K_idx_%d" % k] = np.stack([ii, jj])
        rec["K_val_%d" % k] = gp.K[ii, jj]
        rec["K_diag_%d" % k] = np.diag(gp.K).copy()
        rec["D2_diag_%d" % k] = np.diag(gp.D2).copy()
    np.savez_compressed(os.path.join(HERE, name), **rec)
    print(name, "nlml", [float(rec["nlml_%d" % k]) for k in range(len(thetas))])


if __name__ == "__main__":
    th0 = O.THETA0.copy()
    rng = np.random.default_rng(42)
    th1 = np.clip(th0 * rng.uniform(0.8, 1.25, 10), 1e-4, 6.0)
    th2 = np.clip(th0 * rng.uniform(0.8, 1.25, 10), 1e-4, 6.0)
    make(300, 0, "gp_n300.npz", [th0, th1, th2])
    make(1000, 1, "gp_n1000.npz", [th0, th1])
