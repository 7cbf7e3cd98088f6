import os, sys
import numpy as np, scipy.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
rng = np.random.default_rng(0)
n = 6144
Am = rng.standard_normal((n, n)); S = Am @ Am.T + n * np.eye(n)
Lr = sla.cholesky(S, lower=True)
errs = []
for rep in range(6):
    L, ld, ms, rc = G.test_potrf(S)
    errs.append(np.abs(L - Lr).max() / np.abs(Lr).max())
    if rep == 5:
        bad = np.argwhere(np.abs(L - Lr) > 1e-9 * np.abs(Lr).max())
        if len(bad):
            print("  bad entries", len(bad), "rows", bad[:, 0].min(), bad[:, 0].max(), "cols", bad[:, 1].min(), bad[:, 1].max(), "first", bad[:5].tolist())
            cols = np.unique(bad[:, 1] // 512); print("  bad block cols", cols.tolist())
print("mode", os.environ.get("GPSS_LA_DEBUG"), "errs", ["%.1e" % e for e in errs])
