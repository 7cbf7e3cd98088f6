"""On-GPU check of the opt-in int8 prediction GEMM (GPSS_OZAKI_PREDICT=1) against the default DMMA prediction of the same model.

    GPSS_OZAKI_PREDICT=1 python scripts/oz_predict_check.py [n ...]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen

base = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])
rc = 0
for n, s in [(int(a), 7) for a in sys.argv[1:]] or [(2000, 7), (5000, 7), (20000, 7), (50000, 7)]:
    X, y = datagen.drillholes(n, 0)
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    m_test = 8192 if n >= 20000 else 1000
    Xt = np.concatenate([Xs[:100], np.random.default_rng(1).uniform(-1, 1, (m_test - 100, 3))])
    out = {}
    for pred in (0, 1):
        if n > 8192:
            os.environ.pop("GPSS_OZAKI", None)                    # the library's size rule: 7 slices of 8 bits
            os.environ.pop("GPSS_OZAKI_BITS", None)
        else:
            os.environ["GPSS_OZAKI"] = str(s)
            os.environ["GPSS_OZAKI_BITS"] = "8"
        os.environ["GPSS_OZAKI_PREDICT"] = str(pred)
        m = G.GpssModel(Xs, ys)
        m.set_theta(base)
        m.predict(Xt[:256])                      # warm-up: W, buffers, planes
        t0 = time.perf_counter()
        mu, var = m.predict(Xt)
        out[pred] = (mu, var, m.last_call_ms(), time.perf_counter() - t0)
        m.close()
    e_mu, e_var = np.abs(out[1][0] - out[0][0]).max(), np.abs(out[1][1] - out[0][1]).max()
    ok = e_mu < 1e-8 and e_var < 1e-7
    rc |= 0 if ok else 1
    print("predict n %d S %d m %d: mu %.2e var %.2e | DMMA %.1f ms  int8 %.1f ms (%.0f -> %.0f points/s) %s"
          % (n, s, m_test, e_mu, e_var, out[0][2], out[1][2], m_test / out[0][2] * 1e3, m_test / out[1][2] * 1e3, "ok" if ok else "FAIL"), flush=True)
sys.exit(rc)
