"""On-GPU check of the opt-in int8 tensor-core path (GPSS_OZAKI, gp_ss_ak_b200/csrc/gpss_ozaki.cuh) against the default DMMA
path of the same library: nlml, gradient, alpha, predictive mean / variance at small n, then phase times at large n.

    python scripts/oz_check.py [parity sizes ...] -- [timing sizes ...]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen

args = sys.argv[1:]
split = args.index("--") if "--" in args else len(args)
par_sizes = [int(a) for a in args[:split]] or [2000, 5000]
tim_sizes = [int(a) for a in args[split + 1:]]
S_LIST = [int(s) for s in os.environ.get("OZ_CHECK_S", "8,7").split(",")]
base = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])


def model(Xs, ys, s):
    if s >= 0:
        os.environ["GPSS_OZAKI"] = str(s)             # 0: the FP64 DMMA pipe, 6 | 7 | 8: that many int8 slices
    else:
        os.environ.pop("GPSS_OZAKI", None)            # -1: the library's own size rule
    m = G.GpssModel(Xs, ys)
    os.environ.pop("GPSS_OZAKI", None)
    return m


rc = 0
for n in par_sizes:
    X, y = datagen.drillholes(n, 0)
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    Xt = Xs[:: max(1, n // 50)][:64]
    res = {}
    for s in [0] + S_LIST:
        m = model(Xs, ys, s)
        out = []
        for rep in range(2):                    # second theta: the graph replay / plane reuse path
            th = base * (1 + 0.03 * rep)
            m.set_theta(th)
            L, g = m.nlml_grad()
            a = m.alpha()
            mu, var = m.predict(Xt)
            out.append((L, g, a, mu, var))
        res[s] = out
        m.close()
    for s in S_LIST:
        for rep in range(2):
            L0, g0, a0, mu0, v0 = res[0][rep]
            L, g, a, mu, v = res[s][rep]
            e = (abs(L - L0) / abs(L0), np.abs(g - g0).max() / np.abs(g0).max(), np.abs(a - a0).max() / np.abs(a0).max(),
                 np.abs(mu - mu0).max(), np.abs(v - v0).max())
            ok = e[0] < 1e-9 and e[1] < 1e-7 and e[2] < 1e-8 and e[3] < 1e-8 and e[4] < 1e-7
            rc |= 0 if ok else 1
            print("parity n %d S %d theta %d: nlml rel %.2e  g/max|g| %.2e  alpha %.2e  mu %.2e  var %.2e  %s"
                  % (n, s, rep, e[0], e[1], e[2], e[3], e[4], "ok" if ok else "FAIL"), flush=True)

for n in tim_sizes:
    X, y = datagen.drillholes(n, 0)
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    slist = [int(s) for s in os.environ.get("OZ_TIME_S", "0,7,8").split(",")]
    ref = None
    for s in slist:
        m = model(Xs, ys, s)
        npad = m.padded_n()
        m.set_theta(base)
        L, g = m.nlml_grad()                    # warm-up (allocations)
        m.set_profiling(True)
        m.set_theta(base * 1.01)
        t0 = time.perf_counter()
        L, g = m.nlml_grad()
        wall = time.perf_counter() - t0
        ph = m.phase_ms()
        m.set_profiling(False)
        m.set_theta(base * 1.02)
        m.nlml_grad()
        ms = m.last_call_ms()
        m.set_theta(base * 1.01)
        L, g = m.nlml_grad()
        a = m.alpha()
        if ref is None:
            ref = (L, g, a)
        f3 = float(npad) ** 3 / 3
        print("time n %d S %d bits %s: nlml %.9f (rel to first %.1e, g/max|g| %.1e, alpha %.1e) | eval %.1f ms = %.1f FP64-eq TFLOP/s | potrf %.1f (%.1f) trtri %.1f (%.1f) lauum %.1f (%.1f) "
              "kbuild %.1f solve %.1f grad %.1f" % (n, s, os.environ.get("GPSS_OZAKI_BITS", "7"), L, abs(L - ref[0]) / abs(ref[0]), np.abs(g - ref[1]).max() / np.abs(ref[1]).max(),
                                                   np.linalg.norm(a - ref[2]) / np.linalg.norm(ref[2]), ms,
                                                   3 * f3 / ms * 1e-9, ph[1], f3 / ph[1] * 1e-9, ph[3], f3 / ph[3] * 1e-9, ph[4], f3 / ph[4] * 1e-9,
                                                   ph[0], ph[2], ph[5]), flush=True)
        m.close()
sys.exit(rc)
