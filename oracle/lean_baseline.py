"""TEST / BENCH INFRASTRUCTURE ONLY (like everything under oracle/): the LEAN CPU form of one LML+gradient evaluation that
SURVEY.md section 8(d) asks to be timed beside the reference-literal one.

The reference (GP_Utils.cpp:872-915, 1202-1233) factorises the same matrix three times and forms B^-1 with two n x n triangular
solves: 3 n^3 flop through LAPACK plus ~100 element-wise n x n passes.  A careful CPU implementation of the SAME mathematics needs
one dpotrf (n^3 / 3) and one dpotri (2 n^3 / 3) = n^3 flop and a handful of n^2 passes: that is what this file times, on all host
cores (scipy's bundled OpenBLAS -- the LAPACK real Armadillo would dispatch to).  It is a baseline, not a checker: its numbers are
only sanity-checked against oracle/gpss_oracle.py (tests/test_oracle.py::test_lean_baseline_matches_oracle).
"""
import math
import os
import time

import numpy as np
import scipy.linalg.lapack as lapack


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def lean_eval(Xs, ys, theta, timings=None):
    """nlml, g[6] (Sigma), g[8] (Sigma_Bias), g[9] (sn2) of the reference's objective / gradient with one potrf + one potri.
    (The six rotation / width entries add O(n^2) work only -- gpss_oracle.expans_gradients_fused -- and are left out of the timing
    sample on purpose: the baseline is about the n^3 part.)"""
    from oracle import gpss_oracle as O
    n = Xs.shape[0]
    sn2 = theta[9]
    t0 = time.perf_counter()
    K, D2 = O.compute_K(Xs, Xs, theta, "blas")
    t1 = time.perf_counter()
    A = K + sn2 * np.eye(n)
    c, info = lapack.dpotrf(A, lower=1, overwrite_a=1)
    if info != 0:
        return float("nan"), None
    t2 = time.perf_counter()
    alpha, _ = lapack.dpotrs(c, ys, lower=1)
    nlml = 0.5 * float(ys @ alpha) + float(np.log(np.diag(c)).sum()) + 0.5 * n * math.log(2 * math.pi)
    Ainv, info = lapack.dpotri(c, lower=1, overwrite_c=1)
    t3 = time.perf_counter()
    Ainv = np.tril(Ainv) + np.tril(Ainv, -1).T
    QW = Ainv - np.outer(alpha, alpha)
    E = (K - theta[8]) / theta[6] ** 2                      # exp(-s)
    g6 = 2.0 * theta[6] * float((QW * E).sum())
    g8 = float(np.trace(QW))
    r = ys - K @ alpha
    g9 = -float((Ainv * K).sum()) - float(r @ r) / sn2 + n
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(kbuild=t1 - t0, potrf=t2 - t1, potri=t3 - t2, passes=t4 - t3, total=t4 - t0)
    return nlml, np.array([g6, g8, g9])


def time_lean(sizes, seed=0):
    """[(n, seconds, GFLOP/s of the potrf + potri part)] on all host cores."""
    from gp_ss_ak_b200 import datagen
    from oracle import gpss_oracle as O
    out = []
    for n in sizes:
        X, y = datagen.drillholes(n, seed)
        Xs, ys, _ = datagen.standardise_symmetric(X, y)
        tm = {}
        lean_eval(Xs, ys, O.THETA0, tm)
        out.append((n, tm["total"], float(n) ** 3 / (tm["potrf"] + tm["potri"]) * 1e-9, tm))
    return out


def fit_n2_n3(ns, ts, b_fixed=None):
    """Least-squares t = a n^2 + b n^3 with a, b >= 0 (b fixed when given: the n^3 coefficient of LAPACK work is better known from
    a direct BLAS timing than from small-n samples where the n^2 passes dominate)."""
    ns, ts = np.asarray(ns, float), np.asarray(ts, float)
    if b_fixed is not None:
        a = float(np.sum((ts - b_fixed * ns ** 3) * ns ** 2) / np.sum(ns ** 4))
        return max(a, 0.0), float(b_fixed)
    M = np.stack([ns ** 2, ns ** 3], axis=1)
    from scipy.optimize import nnls
    (a, b), _ = nnls(M, ts)
    return float(a), float(b)


if __name__ == "__main__":
    import sys
    sizes = [int(s) for s in sys.argv[1:]] or [2000, 4000]
    for n, t, gf, tm in time_lean(sizes):
        print("lean n %d: %.2f s (kbuild %.2f potrf %.2f potri %.2f passes %.2f) -> %.0f GFLOP/s on %d threads"
              % (n, t, tm["kbuild"], tm["potrf"], tm["potri"], tm["passes"], gf, blas_threads()))
