import os, sys, time
import numpy as np, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen
from oracle import gpss_oracle as O
rng = np.random.default_rng(0)
for n in (1537, 2048, 2560, 3072, 4096, 6144):
    Am = rng.standard_normal((n, n)); S = Am @ Am.T + n * np.eye(n)
    L, ld, ms, rc = G.test_potrf(S)
    Lr = sla.cholesky(S, lower=True)
    print("potrf n", n, "rc", rc, "err", np.abs(L - Lr).max() / np.abs(Lr).max(), "ms", ms, flush=True)
for n in (3000, 5000, 8000):
    X, y = datagen.drillholes(n, 0)
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    th = O.THETA0.copy()
    Ld, alpha = O.nlml_direct(Xs, ys, th)
    m = G.GpssModel(Xs, ys)
    for rep in range(3):
        m.set_theta(th)
        L, g = m.nlml_grad()
        a = m.alpha()
        print("n", n, "rep", rep, "nlml gpu", L, "direct", Ld, "rel", abs(L - Ld) / abs(Ld), "alpha rel", np.linalg.norm(a - alpha) / np.linalg.norm(alpha), "g6", g[6], flush=True)
    m.close()
