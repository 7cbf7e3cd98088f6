"""CPU numerics study for a ROUND-2 candidate: the O(n^3) contractions of the path (potrf / trtri / lauum bulk updates)
emulated on int8 tensor cores with the Ozaki splitting (error-free slices, int32 accumulation) instead of DMMA.

Nothing here is product code and nothing on the product path uses it.  The script answers one question before any
GPU time is spent: how many 7-bit slices does the path need so that nlml / alpha / g[10] stay inside the parity
tolerances of DESIGN.md section 4 (nlml 1e-9, g 1e-7 of max|g|, alpha 1e-8)?  Integer products are evaluated in
float64, which is exact here (|sum| < 2^53), so the arithmetic is bit-identical to int8 x int8 -> int32 MMA.

    python scripts/ozaki_numerics.py [n] [nb] [row | fixed | kernel]

Slicing (per ROW of each operand, i.e. along k): t = a / 2^e, e = ceil(log2 max|row|); slice l = round-to-nearest of
the running remainder scaled by 2^(7l-1): |q| <= 64, remainder <= 2^-(7l).  Pair (i, j) is kept when i + j <= s + 1
(s (s + 1) / 2 int8 GEMMs); all pairs with the same i + j share one int32 accumulator (exact for k <= 65 536 at s = 8).
"""
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gp_ss_ak_b200 import datagen                    # noqa: E402
from oracle import gpss_oracle as O                   # noqa: E402   (test infrastructure; this is a study script)

BITS = 7


# "fixed" as third argument: ONE a-priori exponent per operand instead of the row maxima (|L_ij| <= sqrt(max B_ii) and
# |W_ij| <= 1 because B = I + K / sn2 >= I) -- what a GPU implementation wants: slices of L / W are then produced once,
# block column by block column, and every product that reads them shares the scale.
# "kernel": the arithmetic of the shipped kernel, operation by operation (oracle/ozaki_oracle.py: a-priori exponent, ONE rounding
# to the 2^-(7S-1) grid, signed base-128 digits, Horner recombination) -- bit-identical to csrc/gpss_ozaki.cuh.
FIXED_SCALE = len(sys.argv) > 3 and sys.argv[3] in ("fixed", "kernel")
KERNEL_ARITH = len(sys.argv) > 3 and sys.argv[3] == "kernel"
FIXED_BOUND = [1.0]


def split_rows(A, s):
    """A (m x k) -> (planes [s, m, k] of integers in [-64, 64], e [m]) with A ~= 2^e sum_l planes[l] 2^-(7 l + 6)."""
    amax = np.abs(A).max(axis=1)
    if FIXED_SCALE:
        amax = np.full_like(amax, FIXED_BOUND[0])
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))), 0.0)
    t = A / np.exp2(e)[:, None]                        # |t| <= 1, exact (power-of-two scaling)
    planes = np.empty((s,) + A.shape)
    rem = t
    for l in range(s):
        sc = np.exp2(BITS * l + BITS - 1)              # 2^6, 2^13, ...
        q = np.rint(rem * sc)
        planes[l] = q
        rem = rem - q / sc                             # exact: q / sc has <= 8 significant bits at that position
    return planes, e


def oz_gemm_nt(A, B, s):
    """A (m x k) @ B (n x k)^T through s slices per operand."""
    if A.shape[1] == 0:
        return np.zeros((A.shape[0], B.shape[0]))
    Ap, ea = split_rows(A, s)
    Bp, eb = split_rows(B, s)
    C = np.zeros((A.shape[0], B.shape[0]))
    for g in range(s - 1, -1, -1):                     # g = i + j (0-based); smallest terms first
        # all pairs of one group in ONE exact integer accumulation: concatenate along k
        Ai = np.concatenate([Ap[i] for i in range(g + 1)], axis=1)
        Bj = np.concatenate([Bp[g - i] for i in range(g + 1)], axis=1)
        P = Ai @ Bj.T
        assert np.abs(P).max() < 2.0 ** 31, "int32 accumulator would overflow"
        C += P * np.exp2(-(BITS * g + 2 * (BITS - 1)))
    return C * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def gemm_nt(A, B, s, bound=1.0):
    FIXED_BOUND[0] = bound
    if s and KERNEL_ARITH:
        # operand kinds as on the GPU: rows of L carry the bound sqrt(max B_ii), rows of U = L^-T the bound 1
        import math
        from oracle import ozaki_oracle as Z
        eL = math.frexp(bound)[1]
        eA = eL if KERNEL_OPERANDS[0] else 0
        eB = eL if KERNEL_OPERANDS[1] else 0
        return Z.oz_gemm_nt(A, B, s, eA, eB)
    return A @ B.T if s == 0 else oz_gemm_nt(A, B, s)


KERNEL_OPERANDS = [True, True]        # which operands of the current product are rows of L (set by the callers below)


def potrf_blocked(Bm, nb, s):
    """Left-looking blocked Cholesky (lower); bulk updates through gemm_nt, diagonal blocks / panel solves in FP64
    (on the GPU those stay on DMMA: O(n^2 nb) flops)."""
    n = Bm.shape[0]
    L = np.tril(Bm).copy()
    bound = float(np.sqrt(np.diag(Bm).max()))          # |L_ij| <= sqrt(B_ii)
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        if j0 > 0:
            KERNEL_OPERANDS[:] = [True, True]
            L[j0:, j0:j1] -= gemm_nt(L[j0:, :j0], L[j0:j1, :j0], s, bound)
        L[j0:j1, j0:j1] = np.linalg.cholesky(np.tril(L[j0:j1, j0:j1]) + np.tril(L[j0:j1, j0:j1], -1).T)
        if j1 < n:
            L[j1:, j0:j1] = sla.solve_triangular(L[j0:j1, j0:j1], L[j1:, j0:j1].T, lower=True).T
    return L


TRTRI_BOUND = [1.0]


def trtri_blocked(L, nb, s):
    """W = L^-1 (lower), block column by block column; bulk product through gemm_nt."""
    n = L.shape[0]
    W = np.zeros_like(L)
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        W[j0:j1, j0:j1] = sla.solve_triangular(L[j0:j1, j0:j1], np.eye(j1 - j0), lower=True)
    for i0 in range(nb, n, nb):                        # block ROW i of W from the rows above it
        i1 = min(i0 + nb, n)
        # W[I, 0:i0] = -W[I,I] (L[I, 0:i0] W[0:i0, 0:i0]) : (nb x i0) @ (i0 x i0), contraction over i0
        KERNEL_OPERANDS[:] = [True, False]
        T = gemm_nt(L[i0:i1, :i0], W[:i0, :i0].T.copy(), s, TRTRI_BOUND[0])
        W[i0:i1, :i0] = -W[i0:i1, i0:i1] @ T
    return W


def lauum_blocked(W, s):
    """B^-1 = W^T W (one contraction over rows, like the product's single lauum launch)."""
    KERNEL_OPERANDS[:] = [False, False]
    return gemm_nt(W.T.copy(), W.T.copy(), s)


def evaluate(X, y, theta, K, D2, nb, s):
    n = X.shape[0]
    sn2 = theta[9]
    Bm = np.eye(n) + K / sn2
    TRTRI_BOUND[0] = float(np.sqrt(np.diag(Bm).max()))
    L = potrf_blocked(Bm, nb, s)
    logdet = float(np.log(np.diag(L)).sum())
    z = sla.solve_triangular(L, y / sn2, lower=True)
    alpha = sla.solve_triangular(L.T, z, lower=False)
    W = trtri_blocked(L, nb, s)
    Q = lauum_blocked(W, s)
    nlml = 0.5 * float(y @ alpha) + logdet + 0.5 * n * np.log(2 * np.pi * sn2)
    QW = Q / sn2 - np.outer(alpha, alpha)
    g = np.zeros(10)
    g[:8] = O.expans_gradients_fused(X, theta, QW, D2)
    g[8] = np.trace(QW)
    r = y - K @ alpha
    g[9] = -float((Q / sn2 * K).sum()) - float(r @ r) / sn2 + n
    return dict(nlml=nlml, logdet=logdet, alpha=alpha, g=g, Q=Q, L=L)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    X, y = datagen.drillholes(n, 0)
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    rng = np.random.default_rng(5)
    thetas = [O.THETA0.copy(), np.clip(O.THETA0 * rng.uniform(0.8, 1.25, 10), 1e-4, 6.0)]
    thetas[1][9] = 1e-3                                # a badly conditioned probe: sn2 16x smaller
    for ti, theta in enumerate(thetas):
        K, D2 = O.compute_K(Xs, Xs, theta)
        cond = np.linalg.cond(np.eye(n) + K / theta[9])
        ref = evaluate(Xs, ys, theta, K, D2, nb, 0)    # same blocked algorithm, plain FP64 products
        lap = evaluate(Xs, ys, theta, K, D2, n, 0)     # unblocked (LAPACK order): the FP64 noise floor
        gs = np.abs(ref["g"]).max()

        def row(tag, r):
            print("  %-10s nlml rel %.2e  logdet rel %.2e  alpha rel %.2e  g/max|g| %.2e  Q rel %.2e" % (
                tag, abs(r["nlml"] - ref["nlml"]) / abs(ref["nlml"]), abs(r["logdet"] - ref["logdet"]) / abs(ref["logdet"]),
                np.abs(r["alpha"] - ref["alpha"]).max() / np.abs(ref["alpha"]).max(), np.abs(r["g"] - ref["g"]).max() / gs,
                np.abs(r["Q"] - ref["Q"]).max() / np.abs(ref["Q"]).max()), flush=True)
        print("theta %d  n %d  nb %d  cond(B) %.2e  nlml %.6f" % (ti, n, nb, cond, ref["nlml"]), flush=True)
        row("fp64-lapack", lap)
        for s in (5, 6, 7, 8, 9):
            r = evaluate(Xs, ys, theta, K, D2, nb, s)
            row("s=%d (%2d)" % (s, s * (s + 1) // 2), r)


if __name__ == "__main__":
    main()
