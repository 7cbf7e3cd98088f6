"""On-GPU timing harness (scratch, not a test): potrf driver and whole-evaluation phase times at several n."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen

sizes = [int(a) for a in sys.argv[1:]] or [2000, 8192, 20000]
base = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])
for n in sizes:
    X, y = datagen.drillholes(n, 0)
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    m = G.GpssModel(Xs, ys)
    npad = m.padded_n()
    m.set_profiling(True)
    for rep in range(3):
        m.set_theta(base * (1 + 0.01 * rep))
        t0 = time.perf_counter()
        L, g = m.nlml_grad()
        wall = time.perf_counter() - t0
        ph = m.phase_ms()
    f3 = float(npad) ** 3 / 3
    print("n %d n_pad %d nlml %.9f wall %.2f ms | kbuild %.2f potrf %.2f (%.2f TF/s) solve %.2f trtri %.2f (%.2f) lauum %.2f (%.2f) grad %.2f | sum %.2f"
          % (n, npad, L, wall * 1e3, ph[0], ph[1], f3 / ph[1] * 1e-9, ph[2], ph[3], f3 / ph[3] * 1e-9, ph[4], f3 / ph[4] * 1e-9, ph[5],
             ph[:6].sum()), flush=True)
    m.set_profiling(False)
    ts = []
    for rep in range(3):
        m.set_theta(base * (1 + 0.02 * rep))
        m.nlml_grad()
        ts.append(m.last_call_ms())
    m.set_theta(base)
    m.nlml()
    t_obj = m.last_call_ms()
    print("   unprofiled nlml_grad ms %s  nlml-only ms %.2f  -> %.3f TF/s on n_pad^3" % (["%.2f" % t for t in ts], t_obj,
                                                                                    float(npad) ** 3 / min(ts) * 1e-9), flush=True)
    m.close()
