"""Profiling target (scratch, not a test): one warm evaluation, then ONE LML+gradient evaluation and one prediction
batch (mean + variance) at the given n, so that ncu captures every kernel of the path once.
   python scripts/profile_target.py [n] [m_test]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m_test = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
base = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])
X, y = datagen.drillholes(n, 0)
Xs, ys, params = datagen.standardise_symmetric(X, y)
m = G.GpssModel(Xs, ys)
m.set_theta(base)
L, g = m.nlml_grad()
Xt_raw, _ = datagen.drillholes(m_test, 5)
Xt = (Xt_raw - params[1:, 0]) / params[1:, 1]
mu, var = m.predict(Xt)
print("n %d n_pad %d nlml %.9f |g|max %.4g  mu[0] %.6f var[1] %.6g  launches %d" % (n, m.padded_n(), L, np.abs(g).max(), mu[0], var[1],
                                                                                  m.launch_count()), flush=True)
m.close()
