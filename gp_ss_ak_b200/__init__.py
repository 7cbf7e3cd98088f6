"""gp_ss_ak_b200 -- B200-native exact-GP hot path of GP_SS_AK.

The product is the shared library libgpss.so (CUDA kernels + C ABI, include/gpss.h) and the C++ host
classes in gp_ss_ak_b200/host that mirror the reference's Kernels / Opt_Algs / GP_utils / gp_ss_ak surface.
This Python package is only the ctypes harness used by tests/ and bench.py: it loads the library and
exposes the C ABI one-to-one.  There is no Python or CPU implementation of the path: if the library is
missing or no CUDA device is present every call raises.
"""
from __future__ import annotations

import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpss.so")
NPAR = 10

GPSS_OK = 0
GPSS_NOT_POSDEF = 1
KERNEL_KINDS = {"ExpAns": 0, "Exp": 1, "RBF": 2}

_lib = None

_c_double_p = ctypes.POINTER(ctypes.c_double)


class GpssError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(_c_double_p) if a is not None else None


def load_library():
    """dlopen libgpss.so and declare the prototypes of every symbol in include/gpss.h (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpssError("libgpss.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for this path)")
    lib = ctypes.CDLL(LIB_PATH)
    H = ctypes.c_void_p
    I, L, D = ctypes.c_int, ctypes.c_long, ctypes.c_double
    P = _c_double_p
    protos = {
        "gpss_version": (I, []),
        "gpss_device_count": (I, [ctypes.POINTER(I)]),
        "gpss_last_error": (ctypes.c_char_p, []),
        "gpss_create": (I, [I, I, I, P, P, ctypes.POINTER(H)]),
        "gpss_destroy": (I, [H]),
        "gpss_set_data": (I, [H, P, P]),
        "gpss_set_theta": (I, [H, P]),
        "gpss_get_theta": (I, [H, P]),
        "gpss_set_kernel": (I, [H, I]),
        "gpss_nlml": (I, [H, P]),
        "gpss_nlml_grad": (I, [H, P, P]),
        "gpss_get_alpha": (I, [H, P]),
        "gpss_get_yhat": (I, [H, P]),
        "gpss_nccl_unique_id": (I, [ctypes.c_void_p]),
        "gpss_dist_init": (I, [H, I, I, ctypes.c_void_p]),
        "gpss_create_partitioned": (I, [I, I, I, ctypes.c_void_p, I, I, P, P, ctypes.POINTER(H)]),
        "gpss_dist_partition": (I, [I, I, I, ctypes.POINTER(I)]),
        "gpss_dist_uslice_layout": (I, [I, I, I, ctypes.POINTER(L), ctypes.POINTER(I), ctypes.POINTER(L)]),
        "gpss_dist_potrf_schedule": (I, [I, I, I, ctypes.POINTER(I), I, ctypes.POINTER(I)]),
        "gpss_predict": (I, [H, L, P, P, P]),
        "gpss_predict_shard": (I, [H, L, P, L, P, P, P]),
        "gpss_var_postprocess": (I, [L, D, P]),
        "gpss_compute_K": (I, [I, I, P, I, I, P, I, P, P, P]),
        "gpss_expans_gradients": (I, [I, P, I, I, P, P, P]),
        "gpss_set_profiling": (I, [H, I]),
        "gpss_get_phase_ms": (I, [H, P]),
        "gpss_get_last_call_ms": (I, [H, P]),
        "gpss_measure_fp64_peak": (I, [I, P]),
        "gpss_measure_int8_peak": (I, [I, P, P]),
        "gpss_get_launch_count": (I, [H, ctypes.POINTER(L)]),
        "gpss_debug_fetch": (I, [H, I, P, L]),
        "gpss_padded_n": (I, [H, ctypes.POINTER(I)]),
        "gpss_set_white": (I, [H, D, I]),
        "gpss_set_kernel2": (I, [H, I]),
        "gpss_set_theta2": (I, [H, P]),
        "gpss_get_grad2": (I, [H, P]),
        "gpss_get_ozaki": (I, [H, ctypes.POINTER(I)]),
        "gpss_get_ozaki_fallbacks": (I, [H, ctypes.POINTER(L)]),
        "gpss_get_ozaki_bits": (I, [H, ctypes.POINTER(I)]),
        "gpss_test_gemm_nt": (I, [I, I, I, I, I, P, P, P, I, P]),
        "gpss_test_oz_gemm": (I, [I, I, I, I, I, P, P, P, I, P]),
        "gpss_test_potrf": (I, [I, I, P, P, P]),
    }
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    lib._gpss_protos = protos
    _lib = lib
    return lib


def exported_symbols():
    return sorted(load_library()._gpss_protos.keys())


def _check(rc, allow_not_posdef=False):
    if rc == GPSS_OK or (allow_not_posdef and rc == GPSS_NOT_POSDEF):
        return rc
    msg = load_library().gpss_last_error()
    raise GpssError("gpss call failed with status %d: %s" % (rc, msg.decode() if msg else ""))


def device_count():
    n = ctypes.c_int(0)
    _check(load_library().gpss_device_count(ctypes.byref(n)))
    return n.value


def _colmajor(X):
    return np.asfortranarray(np.asarray(X, dtype=np.float64))


class GpssModel:
    """Thin handle wrapper: one-to-one with the C ABI (see include/gpss.h for the reference citations)."""

    def __init__(self, X, y, device=0, partitioned=None):
        """partitioned=(rank, world, unique_id): collective constructor of a handle whose factor is stored as block
        columns spread over the ranks (gpss_create_partitioned)."""
        lib = load_library()
        X = _colmajor(X)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        if X.ndim != 2 or X.shape[0] != y.shape[0]:
            raise ValueError("X must be (n, d) and y (n,)")
        self.n, self.d = X.shape
        self._h = ctypes.c_void_p()
        if partitioned is None:
            _check(lib.gpss_create(device, self.n, self.d, _dp(X), _dp(y), ctypes.byref(self._h)))
        else:
            rank, world, uid = partitioned
            buf = ctypes.create_string_buffer(bytes(uid), 128)
            _check(lib.gpss_create_partitioned(device, rank, world, ctypes.cast(buf, ctypes.c_void_p), self.n, self.d, _dp(X), _dp(y),
                                               ctypes.byref(self._h)))
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.gpss_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dist_init(self, rank, world, unique_id):
        """Collective: joins this handle to an NCCL communicator of `world` ranks (unique_id: 128 bytes from rank 0)."""
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        _check(self._lib.gpss_dist_init(self._h, rank, world, ctypes.cast(buf, ctypes.c_void_p)))

    def set_data(self, X, y):
        X = _colmajor(X)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        assert X.shape == (self.n, self.d) and y.shape == (self.n,)
        _check(self._lib.gpss_set_data(self._h, _dp(X), _dp(y)))

    def set_kernel(self, name):
        """Main kernel of Hyb{main, Bias}: "ExpAns" (default, 10 parameters), "Exp" (4) or "RBF" (5)."""
        self._kind = KERNEL_KINDS[name]
        _check(self._lib.gpss_set_kernel(self._h, self._kind))

    def npar(self):
        return (10, 4, 5)[getattr(self, "_kind", 0)]

    def set_theta(self, theta):
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1))
        assert th.shape == (self.npar(),)
        full = np.zeros(NPAR)
        full[:th.shape[0]] = th
        _check(self._lib.gpss_set_theta(self._h, _dp(full)))

    def nlml(self):
        out = ctypes.c_double(0.0)
        _check(self._lib.gpss_nlml(self._h, ctypes.byref(out)), allow_not_posdef=True)
        return out.value

    def nlml_grad(self):
        out = ctypes.c_double(0.0)
        g = np.zeros(NPAR)
        _check(self._lib.gpss_nlml_grad(self._h, ctypes.byref(out), _dp(g)), allow_not_posdef=True)
        return out.value, g[:self.npar()]

    def alpha(self):
        a = np.zeros(self.n)
        _check(self._lib.gpss_get_alpha(self._h, _dp(a)), allow_not_posdef=True)
        return a

    def yhat(self):
        a = np.zeros(self.n)
        _check(self._lib.gpss_get_yhat(self._h, _dp(a)), allow_not_posdef=True)
        return a

    def predict(self, Xs, want_var=True):
        Xs = _colmajor(Xs)
        m = Xs.shape[0]
        mu = np.zeros(m)
        var = np.zeros(m) if want_var else None
        _check(self._lib.gpss_predict(self._h, m, _dp(Xs), _dp(mu), _dp(var)))
        return mu, var

    def predict_shard(self, m_total, sums_total, Xs_shard, want_var=True):
        Xs = _colmajor(Xs_shard)
        cnt = Xs.shape[0]
        mu = np.zeros(cnt)
        var = np.zeros(cnt) if want_var else None
        s = np.ascontiguousarray(np.asarray(sums_total, dtype=np.float64))
        _check(self._lib.gpss_predict_shard(self._h, m_total, _dp(s), cnt, _dp(Xs), _dp(mu), _dp(var)))
        return mu, var

    def set_profiling(self, on=True):
        _check(self._lib.gpss_set_profiling(self._h, 1 if on else 0))

    def phase_ms(self):
        ms = np.zeros(16)
        _check(self._lib.gpss_get_phase_ms(self._h, _dp(ms)))
        return ms

    def last_call_ms(self):
        v = ctypes.c_double(0.0)
        _check(self._lib.gpss_get_last_call_ms(self._h, ctypes.byref(v)))
        return v.value

    def launch_count(self):
        v = ctypes.c_long(0)
        _check(self._lib.gpss_get_launch_count(self._h, ctypes.byref(v)))
        return v.value

    def ozaki_slices(self):
        """0 = the long-k contractions run on the FP64 DMMA pipe; 6 | 7 | 8 = on the int8 tensor cores with that many slices."""
        v = ctypes.c_int(0)
        _check(self._lib.gpss_get_ozaki(self._h, ctypes.byref(v)))
        return v.value

    def set_member2(self, kind2, theta2=None):
        """Second distance-based member of the Hyb sum: kind2 = -1 (none) | 0 ExpAns | 1 Exp | 2 RBF, theta2 = its own parameters."""
        _check(self._lib.gpss_set_kernel2(self._h, int(kind2)))
        if theta2 is not None:
            t = np.zeros(8)
            t[:len(theta2)] = theta2
            _check(self._lib.gpss_set_theta2(self._h, _dp(t)))

    def grad2(self):
        """Gradient entries of the second member from the last nlml_grad() (8 slots, the member's own parameter order)."""
        g = np.zeros(8)
        _check(self._lib.gpss_get_grad2(self._h, _dp(g)))
        return g

    def set_white(self, sigma_white, cross_diagonal=False):
        """Sum of the White members' Sigma_White (Kernel.cpp:180-270); cross_diagonal: see include/gpss.h."""
        _check(self._lib.gpss_set_white(self._h, float(sigma_white), 1 if cross_diagonal else 0))

    def ozaki_digit_bits(self):
        v = ctypes.c_int(0)
        _check(self._lib.gpss_get_ozaki_bits(self._h, ctypes.byref(v)))
        return v.value

    def ozaki_fallbacks(self):
        """Evaluations repeated on the DMMA pipe because an int8 operand left its a-priori bound (device flag)."""
        v = ctypes.c_long(0)
        _check(self._lib.gpss_get_ozaki_fallbacks(self._h, ctypes.byref(v)))
        return v.value

    def padded_n(self):
        v = ctypes.c_int(0)
        _check(self._lib.gpss_padded_n(self._h, ctypes.byref(v)))
        return v.value

    def debug_fetch(self, which):
        """0 = L factor, 1 = U = L^-T, 2 = B^-1 / L^-1 (n_pad x n_pad each); 3 = the transformed coordinates (n_pad x 5)."""
        npad = self.padded_n()
        out = np.zeros((npad, 5 if which == 3 else npad), order="F")
        _check(self._lib.gpss_debug_fetch(self._h, which, _dp(out), out.size))
        return out


def var_postprocess(var_raw, sn2):
    """The reference's post-processing of the gathered raw variance vector (include/gpss.h)."""
    v = np.ascontiguousarray(np.asarray(var_raw, dtype=np.float64).reshape(-1)).copy()
    _check(load_library().gpss_var_postprocess(v.shape[0], float(sn2), _dp(v)))
    return v


def nccl_unique_id():
    buf = ctypes.create_string_buffer(128)
    _check(load_library().gpss_nccl_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
    return bytes(buf.raw)


def dist_partition(n_pad, world, kind):
    b = (ctypes.c_int * (world + 1))()
    _check(load_library().gpss_dist_partition(n_pad, world, kind, b))
    return list(b)


def dist_uslice_layout(n_pad, r0, rows):
    """(offsets, lens, count) of the packed exchange layout of the U = L^-T row slice [r0, r0 + rows) (host logic only)."""
    ncol = n_pad - r0
    off = (ctypes.c_long * ncol)()
    lens = (ctypes.c_int * ncol)()
    cnt = ctypes.c_long(0)
    _check(load_library().gpss_dist_uslice_layout(n_pad, r0, rows, off, lens, ctypes.byref(cnt)))
    return list(off), list(lens), cnt.value


def dist_potrf_schedule(nblk, world, rank):
    """[(kind, col, first_panel, panel_count, root, stream), ...] executed by `rank` (host logic only)."""
    cnt = ctypes.c_int(0)
    lib = load_library()
    _check(lib.gpss_dist_potrf_schedule(nblk, world, rank, None, 0, ctypes.byref(cnt)))
    buf = (ctypes.c_int * (6 * cnt.value))()
    _check(lib.gpss_dist_potrf_schedule(nblk, world, rank, buf, cnt.value, ctypes.byref(cnt)))
    return [tuple(buf[6 * i:6 * i + 6]) for i in range(cnt.value)]


def measure_fp64_peak(device=0):
    v = ctypes.c_double(0.0)
    _check(load_library().gpss_measure_fp64_peak(device, ctypes.byref(v)))
    return v.value


def measure_int8_peak(device=0):
    """(burst, sustained) int8 tensor-pipe micro-peak in TOP/s (tcgen05 kind::i8, M 128 N 256 K 32, operands resident in shared memory)."""
    b, s = ctypes.c_double(0.0), ctypes.c_double(0.0)
    _check(load_library().gpss_measure_int8_peak(device, ctypes.byref(b), ctypes.byref(s)))
    return b.value, s.value


def compute_K(theta, X1, X2, want_K=True, want_D2=True, device=0):
    """K and D2 of Hyb{main, Bias}; the main kernel follows from len(theta): 10 ExpAns, 4 Exp, 5 RBF."""
    lib = load_library()
    X1 = _colmajor(X1)
    X2 = _colmajor(X2)
    th0 = np.asarray(theta, dtype=np.float64).reshape(-1)
    kind = {10: 0, 4: 1, 5: 2}[th0.shape[0]]
    th = np.zeros(NPAR)
    th[:th0.shape[0]] = th0
    K = np.zeros((X1.shape[0], X2.shape[0]), order="F") if want_K else None
    D2 = np.zeros((X1.shape[0], X2.shape[0]), order="F") if want_D2 else None
    _check(lib.gpss_compute_K(device, kind, _dp(th), X1.shape[1], X1.shape[0], _dp(X1), X2.shape[0], _dp(X2), _dp(K), _dp(D2)))
    return K, D2


def expans_gradients(theta, X, QW, device=0):
    """Kern_ExpAnisotropic::getGradients for a host QW (compatibility entry point); returns g[0..7]."""
    X = _colmajor(X)
    QW = _colmajor(QW)
    th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
    g = np.zeros(8)
    _check(load_library().gpss_expans_gradients(device, _dp(th), X.shape[1], X.shape[0], _dp(X), _dp(QW), _dp(g)))
    return g


def test_gemm_nt(A, B, C=None, tile=0, device=0):
    """C = A @ B.T (or C - A @ B.T when C is given) through the DMMA kernel; returns (C, ms)."""
    lib = load_library()
    A = _colmajor(A)
    B = _colmajor(B)
    M, K = A.shape
    N = B.shape[0]
    sub = C is not None
    Cc = _colmajor(C).copy(order="F") if sub else np.zeros((M, N), order="F")
    ms = ctypes.c_double(0.0)
    _check(lib.gpss_test_gemm_nt(device, tile, M, N, K, _dp(A), _dp(B), _dp(Cc), 1 if sub else 0, ctypes.byref(ms)))
    return Cc, ms.value


def test_oz_gemm(A, B, C=None, slices=8, device=0):
    """C = A @ B.T (or C - A @ B.T when C is given) through the int8 tensor-core kernel; |A|, |B| <= 1; returns (C, ms)."""
    lib = load_library()
    A = _colmajor(A)
    B = _colmajor(B)
    M, K = A.shape
    N = B.shape[0]
    sub = C is not None
    Cc = _colmajor(C).copy(order="F") if sub else np.zeros((M, N), order="F")
    ms = ctypes.c_double(0.0)
    _check(lib.gpss_test_oz_gemm(device, slices, M, N, K, _dp(A), _dp(B), _dp(Cc), 1 if sub else 0, ctypes.byref(ms)))
    return Cc, ms.value


def test_potrf(A, device=0):
    """Blocked Cholesky (lower) of an SPD matrix through the potrf driver; returns (L, sum(log(diag L)), ms, status)."""
    lib = load_library()
    Ac = _colmajor(A).copy(order="F")
    n = Ac.shape[0]
    ld = ctypes.c_double(0.0)
    ms = ctypes.c_double(0.0)
    rc = lib.gpss_test_potrf(device, n, _dp(Ac), ctypes.byref(ld), ctypes.byref(ms))
    _check(rc, allow_not_posdef=True)
    return Ac, ld.value, ms.value, rc
