"""CPU tests of the numpy restatement of the int8 tensor-core product (oracle/ozaki_oracle.py restates
gp_ss_ak_b200/csrc/gpss_ozaki.cuh operation by operation): digit arithmetic, exactness of the integer accumulation, error
bounds of the FP64 result, and the numerics of the whole blocked path (which products run through it) against the oracle's
tolerances.  The GPU side of the same statement is tests/test_gpu_parity.py::test_int8_*."""
import math
from fractions import Fraction

import numpy as np
import pytest
import scipy.linalg as sla

from gp_ss_ak_b200 import datagen
from oracle import gpss_oracle as O
from oracle import ozaki_oracle as Z


@pytest.mark.parametrize("S", [6, 7, 8])
def test_digits_are_an_exact_fixed_point_expansion(S):
    rng = np.random.default_rng(S)
    e = 3
    x = np.concatenate([rng.uniform(-8, 8, 4000) * 2.0 ** -rng.integers(0, 40, 4000), [8.0, -8.0, 0.0, 7.999999999999999, 2.0 ** -60]])
    d = Z.oz_digits(x.reshape(1, -1), e, S)
    assert d.min() >= -64 and d.max() <= 64                      # int8 range, top digit reaches +-64 only at |x| = 2^e
    assert np.all(d[1:] <= 63)
    # the digits reproduce v = rint(x 2^(7S-1-e)) exactly (python integers: no rounding anywhere)
    for k in range(x.size):
        v = sum(int(d[p, 0, k]) * 128 ** (S - 1 - p) for p in range(S))
        exact = Fraction(float(x[k])) * Fraction(2) ** (7 * S - 1 - e)
        assert abs(Fraction(v) - exact) <= Fraction(1, 2)
    # and stand for x to half a unit of the grid 2^(e-7S+1)
    if S <= 7:
        assert np.abs(Z.oz_undigits(d, e)[0] - x).max() <= 2.0 ** (e - 7 * S)


@pytest.mark.parametrize("S", [6, 7])
def test_eight_bit_digits_use_the_full_int8_range(S):
    """GPSS_OZAKI_BITS=8: base-256 signed digits; 7 of them carry the bits that take 8 digits of 7 bits.  The top digit fits int8
    only below 127.5/128 of the scale, which the widened exponent guarantees for every value up to the operand's bound."""
    rng = np.random.default_rng(S)
    e = Z.oz_exponent(Z.SCALE_UNIT, bits=8)
    assert e == 1
    x = np.concatenate([rng.uniform(-1, 1, 4000) * 2.0 ** -rng.integers(0, 40, 4000), [1.0, -1.0, 0.0, 0.9999999999999999]])
    d = Z.oz_digits(x.reshape(1, -1), e, S, bits=8)
    assert d.min() >= -128 and d.max() <= 127
    for k in range(x.size):
        v = sum(int(d[p, 0, k]) * 256 ** (S - 1 - p) for p in range(S))
        exact = Fraction(float(x[k])) * Fraction(2) ** (8 * S - 1 - e)
        assert abs(Fraction(v) - exact) <= Fraction(1, 2)
    # a Cholesky-kind bound just below a power of two takes the next exponent, one well below it does not
    assert Z.oz_exponent(Z.SCALE_CHOL, sigma2=62.9 * 0.016, bias=0.0, sn2=0.016, bits=8) == Z.oz_exponent(Z.SCALE_CHOL, sigma2=62.9 * 0.016, bias=0.0, sn2=0.016) + 1
    assert Z.oz_exponent(Z.SCALE_CHOL, theta=O.THETA0, bits=8) == Z.oz_exponent(Z.SCALE_CHOL, theta=O.THETA0)
    # values that would break the bound saturate at [127, ..., 127] instead of wrapping
    d = Z.oz_digits(np.array([[3.0, -3.0]]), 0, S, bits=8)
    assert d.min() >= -128 and d.max() <= 127
    A = rng.uniform(-1, 1, (24, 40000))
    B = rng.uniform(-1, 1, (16, 40000))
    assert Z.oz_kseg(S, 8) % 64 == 0 and S * Z.oz_kseg(S, 8) * 2 ** 14 < 2 ** 31
    C = Z.oz_gemm_nt(A, B, S, e, e, bits=8)                          # three k-segments, each exact in int32 (asserted inside)
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    assert np.abs(C - ref).max() <= 3 * 40000 * 2.0 ** (2 * e - (8 * S - 1)) + 4 * 2.0 ** -53 * np.abs(ref).max()


def test_values_beyond_the_bound_are_clamped_not_wrapped():
    d = Z.oz_digits(np.array([[1.0000001, -3.0]]), 0, 7)
    assert d.min() >= -64 and d.max() <= 64
    assert np.array_equal(Z.oz_undigits(d, 0)[0], [1.0, -1.0])


@pytest.mark.parametrize("S,k", [(7, 512), (8, 1024)])
def test_product_error_bound_and_int32_range(S, k):
    rng = np.random.default_rng(11)
    A = rng.uniform(-1, 1, (48, k)) * 2.0 ** -rng.integers(0, 12, (48, k))
    B = rng.uniform(-1, 1, (40, k)) * 2.0 ** -rng.integers(0, 12, (40, k))
    C = Z.oz_gemm_nt(A, B, S)
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    # each operand is off by <= 2^-(7S) (half a grid unit), every dropped pair i + j >= S is below 2^-(7S-1) per term
    assert np.abs(C - ref).max() <= 3 * k * 2.0 ** -(7 * S - 1)
    # worst-case digits still fit the int32 accumulators at the sizes the library admits (S n_pad 4096 < 2^31)
    big = np.full((1, 57344), 1.0)
    G = Z.oz_groups(Z.oz_digits(big, 0, 8), Z.oz_digits(big, 0, 8))
    assert max(int(np.abs(g).max()) for g in G) < 2 ** 31


def test_eight_slices_beat_plain_fp64_accumulation():
    """S = 8 rounds each operand once at 2^-55 of its bound and accumulates EXACTLY; a chain of k FP64 FMAs does not."""
    rng = np.random.default_rng(3)
    k = 4096
    A = rng.uniform(-1, 1, (32, k))
    B = rng.uniform(-1, 1, (32, k))
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    e_oz = np.abs(Z.oz_gemm_nt(A, B, 8) - ref).max()
    seq = np.zeros((32, 32))
    for q in range(k):                                          # the order a single accumulator sees
        seq += np.outer(A[:, q], B[:, q])
    assert e_oz <= np.abs(seq - ref).max()


def test_subtract_form_and_scales():
    rng = np.random.default_rng(5)
    A = rng.uniform(-7, 7, (16, 128))
    B = rng.uniform(-0.9, 0.9, (24, 128))
    C0 = rng.uniform(-1, 1, (16, 24))
    out = Z.oz_gemm_nt(A, B, 8, eA=3, eB=0, C=C0, sign=-1.0)
    assert np.abs(out - (C0 - A @ B.T)).max() <= 128 * 8 * 2.0 ** -54
    assert Z.oz_exponent(Z.SCALE_UNIT) == 0
    th = O.THETA0
    e = Z.oz_exponent(Z.SCALE_CHOL, theta=th)
    bii = 1.0 + (th[6] ** 2 + th[8]) / th[9]
    assert 2.0 ** (e - 1) <= math.sqrt(bii) < 2.0 ** e


@pytest.mark.parametrize("S,bits,ok", [(8, 7, True), (7, 7, True), (4, 7, False), (7, 8, True), (6, 8, True)])
def test_blocked_path_numerics_against_oracle_tolerances(S, bits, ok):
    """The device path's structure (long-k updates of potrf / trtri / lauum through the int8 product with a-priori scales,
    diagonal blocks and panel solves in FP64) at n = 600: S = 7 and 8 stay inside the parity tolerances of
    tests/test_gpu_parity.py, S = 4 does not (the check is not vacuous)."""
    n, nb = 600, 128
    X, y = datagen.drillholes(n, 0)
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    th = O.THETA0.copy()
    K, D2 = O.compute_K(Xs, Xs, th)
    sn2 = th[9]
    Bm = np.eye(n) + K / sn2
    eL, eU = Z.oz_exponent(Z.SCALE_CHOL, theta=th, bits=bits), Z.oz_exponent(Z.SCALE_UNIT, bits=bits)
    assert np.sqrt(np.diag(Bm).max()) < 2.0 ** eL

    def prod(A, B, ka, kb):
        return Z.oz_gemm_nt(A, B, S, eL if ka == Z.SCALE_CHOL else eU, eL if kb == Z.SCALE_CHOL else eU, bits=bits)

    def evaluate(p):
        L = Z.potrf_blocked(Bm, nb, p)
        alpha = sla.solve_triangular(L.T, sla.solve_triangular(L, ys / sn2, lower=True), lower=False)
        U = Z.trtri_blocked(L, nb, p)
        assert np.abs(U).max() <= 1.0                            # the a-priori bound the UNIT scale rests on (B >= I)
        Q = np.tril(Z.lauum(U, p))
        Q = Q + np.tril(Q, -1).T
        nlml = 0.5 * float(ys @ alpha) + float(np.log(np.diag(L)).sum()) + 0.5 * n * math.log(2 * math.pi * sn2)
        g = O.expans_gradients_fused(Xs, th, Q / sn2 - np.outer(alpha, alpha), D2)
        return nlml, alpha, g

    L0, a0, g0 = evaluate(lambda A, B, ka, kb: A @ B.T)
    L1, a1, g1 = evaluate(prod)
    errs = (abs(L1 - L0) / abs(L0), np.linalg.norm(a1 - a0) / np.linalg.norm(a0), np.abs(g1 - g0).max() / np.abs(g0).max())
    inside = errs[0] <= 1e-9 and errs[1] <= 1e-8 and errs[2] <= 1e-7
    assert inside == ok, errs


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_digit_plane_coverage_of_the_row_sliced_inverse(world):
    """Host logic of the int8 path in trtri_upper / lauum_lower (csrc/gpss_inverse.cuh), replayed for every rank: each k-chunk of
    U planes a product reads must have been cut before, from rows the rank owns or computes redundantly (the diagonal blocks).
    world = 1 is the shipped single-GPU path, world > 1 the opt-in GPSS_OZAKI_DIST layout."""
    import gp_ss_ak_b200 as G
    NBO, BM = 512, 128
    n_pad = 20096                                                    # 39 full block columns + a 128-wide remainder
    nblk_o = (n_pad + NBO - 1) // NBO
    ub = G.dist_partition(n_pad, world, 0) if world > 1 else [0, n_pad]
    qb = G.dist_partition(n_pad, world, 1) if world > 1 else [0, n_pad]
    for r in range(world):
        R0, R1 = ub[r], ub[r + 1]
        assert R0 % BM == 0 and R1 % BM == 0
        cut = np.zeros((n_pad // BM, nblk_o), dtype=bool)            # [row tile, block column of k] of the U planes on this rank
        for t in range(nblk_o):
            J0 = t * NBO
            nbj = min(NBO, n_pad - J0)
            sr0, sr1 = R0, min(R1, J0 + nbj)
            ra, rb = R0, min(R1, J0)
            if t > 0 and rb > ra:
                # product (3): rows [ra, rb), k from each row tile's own first row up to J0
                for rt in range(ra // BM, rb // BM):
                    for kb in range((rt * BM) // NBO, t):
                        assert cut[rt, kb], "rank %d step %d reads uncut planes (row tile %d, block column %d)" % (r, t, rt, kb)
            if sr1 > sr0:
                cut[sr0 // BM:(sr1 + BM - 1) // BM, t] = True        # slice of block column t after (4) / of the diagonal block
        # B^-1 rows [q0, q1): distributed handles re-cut every row below q1 after the all-gather
        q0, q1 = qb[r], qb[r + 1]
        if world > 1:
            cut[:q1 // BM, :] = True
        for rt in range(q0 // BM, q1 // BM):
            for kb in range((rt * BM) // NBO, nblk_o):
                assert cut[rt, kb]                                   # A rows
        # B rows 0 .. q1 meet the same k-range as the A tile they are paired with (k >= the A tile's first row >= their own)
        for rt in range(0, q1 // BM):
            for kb in range((max(rt, q0 // BM) * BM) // NBO, nblk_o):
                assert cut[rt, kb]


def test_leading_digits_are_the_shorter_expansion():
    """GPSS_OZAKI_GRAD: a product may read only the top planes of a tensor cut with more slices -- the leading 7 digits of the
    8-digit signed expansion stand for the value rounded to 7 digits (to within the second rounding)."""
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, 5000) * 2.0 ** -rng.integers(0, 30, 5000)
    d8 = Z.oz_digits(x.reshape(1, -1), 0, 8)
    top7 = Z.oz_undigits(d8[:7], 0)[0]
    assert np.abs(top7 - x).max() <= (0.5 + 1.0 / 128) * 2.0 ** -48
    d7 = Z.oz_digits(x.reshape(1, -1), 0, 7)
    assert (np.abs(d8[:7] - d7).sum(axis=0) != 0).mean() < 0.02          # they differ only where the 8th digit sits on a tie


def test_bench_int8_op_count_mirrors_the_kernel_tiles():
    """bench.py's roofline numerator: the MACs per slice pair the three int8 launches of an evaluation execute, tile by tile."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    def brute(n_pad, nbo=512):
        mac, T = 0, (n_pad + nbo - 1) // nbo
        for t in range(1, T - 1):                                   # oz launch of U1(t + 1): every 128 x 64 tile not above the diagonal
            T0, T1 = t * nbo, (t + 1) * nbo
            nb1 = min(nbo, n_pad - T1)
            for tm in range((n_pad - T1) // 128):
                for tn in range(nb1 // 64):
                    if T1 + tn * 64 > T1 + tm * 128 + 127:
                        continue
                    mac += 128 * 64 * T0
        for t in range(1, T):                                       # product (3) of the inverse: k from the tile's first row to J0
            J0 = t * nbo
            nbj = min(nbo, n_pad - J0)
            for tm in range(J0 // 128):
                for tn in range(nbj // 64):
                    mac += 128 * 64 * (J0 - tm * 128)
        for tm in range(n_pad // 128):                              # B^-1: lower tiles, k from the tile's first row to n_pad
            for tn in range(n_pad // 64):
                if tn * 64 > tm * 128 + 127:
                    continue
                mac += 128 * 64 * (n_pad - tm * 128)
        return mac

    for n_pad in (1024, 2176, 5120):
        assert bench.int8_macs(n_pad) == brute(n_pad)
    for n_pad in (20096, 50048):
        share = 2.0 * bench.int8_macs(n_pad) / float(n_pad) ** 3
        assert 0.95 < share < 1.0
