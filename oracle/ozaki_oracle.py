"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): numpy restatement of the int8 tensor-core evaluation of the long-k
FP64 contractions (gp_ss_ak_b200/csrc/gpss_ozaki.cuh), operation by operation, so that the CUDA kernel can be checked BIT-EXACTLY:

  * oz_digits      = oz_slice_kernel: v = rint(x 2^(7S-1-e)) (round-half-even, clamped), signed base-128 digits in [-64, 64];
  * oz_gemm_nt     = oz_gemm_kernel<S>: G_g = sum_{i+j=g} A_i B_j^T in exact integers (asserted to fit int32 like the TMEM
                     accumulators), Horner recombination in FP64 from the smallest group, C (+)= sign 2^(eA+eB-12) acc.
    Every FP64 operation of the epilogue other than the Horner additions is a multiplication by a power of two, so fused and
    unfused evaluation agree and the restatement reproduces the device result to the last bit.
  * oz_exponent    = oz_exponent() of the kernel file: the a-priori power-of-two bound of an operand kind.
  * blocked potrf / trtri / lauum with a pluggable product: the structure of the device path (which products go through the
    int8 kernel, which stay FP64), used by the CPU numerics study scripts/ozaki_numerics.py and by tests/test_ozaki_cpu.py.

The reference (Armadillo / LAPACK) has no counterpart of this file: it restates OUR kernel, not the reference; the reference
parity of the path that contains the kernel is established by oracle/gpss_oracle.py and the golden fixtures.
"""
import math

import numpy as np
import scipy.linalg as sla

DIGIT_BITS = 7
SCALE_UNIT, SCALE_CHOL = 0, 1


def oz_exponent(kind, theta=None, sigma2=None, bias=None, sn2=None, bits=DIGIT_BITS):
    """e with |x| < 2^e for every element of the operand: 0 for U = L^-T (B >= I), frexp exponent of sqrt(B_ii) for L.
    8-bit digits: the bound is widened by 128 / 127.5 first, so that the top base-256 digit stays <= 127."""
    bound = 1.0
    if kind != SCALE_UNIT:
        if theta is not None:
            sigma2, bias, sn2 = theta[6] ** 2, theta[8], theta[9]
        sw = math.sqrt(1.0 / sn2)
        sww = sw * sw                                    # DevParams.sww = fl(Sw * Sw) (gpss_ctx.cuh fill_params)
        bound = math.sqrt(1.0 + sww * (sigma2 + bias))
    elif bits != 8:
        return 0
    if bits == 8:
        bound *= 128.0 / 127.5
    return math.frexp(bound)[1]


def oz_digits(X, e, S, bits=DIGIT_BITS):
    """[S, rows, k] int64 digits d_p with  x ~= 2^e sum_p d_p 2^-(b p + b - 1)  (one rounding, at 2^(e - b S + 1)).
    bits = 7 (shipped default): digits in [-64, 64].  bits = 8 (GPSS_OZAKI_BITS=8): the full int8 range [-128, 127]; the top
    digit fits only for |v| < 127.5 x 256^(S-1), which oz_exponent(bits=8) guarantees (v is also saturated at [127, ..., 127])."""
    lim = float(1 << (bits * S - 1))
    v = np.rint(np.clip(np.asarray(X, dtype=np.float64) * math.ldexp(lim, -e), -lim, lim)).astype(np.int64)
    vmax = 127 * (((1 << (8 * min(S, 7))) - 1) // 255) if bits == 8 else 1 << (DIGIT_BITS * S - 1)
    v = np.clip(v, -vmax, vmax)
    half, mask = 1 << (bits - 1), (1 << bits) - 1
    d = np.empty((S,) + v.shape, dtype=np.int64)
    for p in range(S - 1, 0, -1):
        dg = ((v + half) & mask) - half
        v = (v - dg) >> bits
        d[p] = dg
    d[0] = v
    return d


def oz_undigits(d, e, bits=DIGIT_BITS):
    """The value the digits stand for (exact in FP64 while bits * S <= 53; for tests)."""
    S = d.shape[0]
    v = np.zeros(d.shape[1:], dtype=np.int64)
    for p in range(S):
        v = v * (1 << bits) + d[p]
    return v.astype(np.float64) * math.ldexp(1.0, e - (bits * S - 1))


def oz_groups(Ad, Bd):
    """G_g = sum_{i+j=g} A_i B_j^T, g < S, as exact int64; asserts the int32 range of the TMEM accumulators."""
    S = Ad.shape[0]
    Af, Bf = Ad.astype(np.float64), Bd.astype(np.float64)     # BLAS on integer-valued doubles: exact, |sums| < 2^31 << 2^53
    out = []
    for g in range(S):
        G = np.zeros((Ad.shape[1], Bd.shape[1]))
        for i in range(g + 1):
            G += Af[i] @ Bf[g - i].T
        assert np.abs(G).max(initial=0) < 2 ** 31, "int32 accumulator overflow"
        out.append(G.astype(np.int64))
    return out


def oz_kseg(S, bits=DIGIT_BITS):
    """Longest k-range one int32 accumulation may cover: |G_g| <= S k 2^(2 (bits - 1)) < 2^31, rounded down to a multiple of 64."""
    return ((1 << (31 - 2 * (bits - 1))) // S - 1) // 64 * 64 if bits == 8 else 1 << 30


def oz_gemm_nt(A, B, S, eA=0, eB=0, C=None, sign=1.0, bits=DIGIT_BITS):
    """sign * A B^T (C given: C + sign * A B^T) exactly as oz_gemm_kernel<S> evaluates it.  With 8-bit digits the k-range is cut
    into segments of oz_kseg (one launch each, the later ones accumulating into C in FP64), as oz_gemm_on does."""
    Ad, Bd = oz_digits(A, eA, S, bits), oz_digits(B, eB, S, bits)
    w = math.ldexp(1.0, -bits)
    scale = sign * math.ldexp(1.0, eA + eB - 2 * (bits - 1))
    out = None if C is None else np.array(C, dtype=np.float64)
    kseg = oz_kseg(S, bits)
    for k0 in range(0, Ad.shape[2], kseg):
        G = oz_groups(Ad[:, :, k0:k0 + kseg], Bd[:, :, k0:k0 + kseg])
        acc = np.zeros(G[0].shape)
        for g in range(S - 1, -1, -1):
            acc = acc * w + G[g].astype(np.float64)
        out = scale * acc if out is None else out + scale * acc
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the blocked path with a pluggable long-k product (structure of gpss_potrf.cuh / gpss_inverse.cuh on one GPU)
# ---------------------------------------------------------------------------------------------------------------------
def potrf_blocked(Bm, nb, prod):
    """Left-looking blocked Cholesky (lower).  prod(A, B, kindA, kindB) -> A B^T evaluates the bulk updates; diagonal blocks
    and panel solves stay FP64 (on the GPU: the DMMA panel work)."""
    n = Bm.shape[0]
    L = np.tril(Bm).copy()
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        if j0 > 0:
            L[j0:, j0:j1] -= prod(L[j0:, :j0], L[j0:j1, :j0], SCALE_CHOL, SCALE_CHOL)
        L[j0:j1, j0:j1] = np.linalg.cholesky(np.tril(L[j0:j1, j0:j1]) + np.tril(L[j0:j1, j0:j1], -1).T)
        if j1 < n:
            L[j1:, j0:j1] = sla.solve_triangular(L[j0:j1, j0:j1], L[j1:, j0:j1].T, lower=True).T
    return L


def trtri_blocked(L, nb, prod):
    """U = L^-T (upper), block column by block column: U[0:J, J] = -(U[0:J, 0:J] L[J, 0:J]^T) U[J, J] (trtri_upper)."""
    n = L.shape[0]
    U = np.zeros_like(L)
    for j0 in range(0, n, nb):
        j1 = min(j0 + nb, n)
        Ujj = sla.solve_triangular(L[j0:j1, j0:j1], np.eye(j1 - j0), lower=True).T
        U[j0:j1, j0:j1] = Ujj
        if j0 > 0:
            T = prod(U[:j0, :j0], L[j0:j1, :j0], SCALE_UNIT, SCALE_CHOL)
            U[:j0, j0:j1] = -T @ Ujj
    return U


def lauum(U, prod):
    """B^-1 = U U^T (lauum_lower; one launch on the GPU)."""
    return prod(U, U, SCALE_UNIT, SCALE_UNIT)
