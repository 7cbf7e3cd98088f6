"""CPU oracle for the exact-GP hot path of GP_SS_AK  --  TEST INFRASTRUCTURE ONLY.

This file restates, step by step, the arithmetic the reference performs on the path
named by BASELINE.json:north_star.  It is imported only by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  The
product (gp_ss_ak_b200/) never imports it and has no CPU fallback.

PARITY PINNED AGAINST THE REFERENCE ITSELF.  The reference ships no tests, golden vectors or
fixtures and real Armadillo is absent from this image, but its sources
compile UNMODIFIED against the in-repo Armadillo stand-in (gp_ss_ak_b200/host/armadillo_standin,
dense kernels forwarded to scipy's OpenBLAS): oracle/Makefile builds oracle/_ref/gp_ss_ak and
oracle/_ref/ref_driver from the files where they lie under /root/reference, and
tests/golden/make_ref_golden.py stores what the reference's own classes compute
(tests/golden/ref_n300.npz, ref_n1000.npz: standardisation, nlml, g[10], Alpha, K samples,
predictions incl. the variance post-processing quirk, a 30-iteration LBFGS probe trace).
tests/test_oracle.py::test_oracle_matches_compiled_reference holds this restatement to those
numbers: standardisation bit-exact; nlml 2e-7, g 5e-7, alpha 5e-7 relative, mu 5e-7 / var 1e-7
absolute.  Those are NOT tolerances of the restatement but the reference's own reproducibility
floor: sqrt() of the O(1e-16) rounding residue that MahaDist's expansion-form distance leaves on
~5% of the diagonal of D2 moves K_ii by ~3e-8, and which entries carry a residue depends on the
BLAS dgemm micro-kernel (hard part 1).  Against this oracle's own
defined-operation-order distance the CUDA path is held to 1e-9 (tests/test_gpu_parity.py).

All file:line citations are into /root/reference.

LAPACK/BLAS: numpy / scipy dispatch to the bundled OpenBLAS (dpotrf, dtrtrs, dgemm,
dgemv) -- the same routines Armadillo forwards chol(), solve(trimatl/trimatu()), and
operator* to.
"""
from __future__ import annotations

import math
import numpy as np
import scipy.linalg as sla

# ----------------------------------------------------------------------------------
# parameter vector (GP_Utils.cpp:101-128, Kernel.cpp:741-773, Kernel.cpp:317-320)
#   theta = [AngleX, iWx, AngleY, iWy, AngleZ, iWz, Sigma, iWR, Sigma_Bias, sn2]
# ----------------------------------------------------------------------------------
THETA0 = np.array([math.pi / 3.1, 1.5, math.pi / 3.1, 1.5, math.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])
NPAR = 10
# The isotropic members of the kernel family.  A parameter vector is always
# [kernel parameters..., Sigma_Bias, sn2]; its LENGTH names the main kernel of the Hyb{main, Bias} covariance:
#   10  ExpAns  (8 kernel parameters, Kernel.cpp:763-838)
#    4  Exp     (Hayper_Euc_Exp, Sigma_Exp; Kernel.cpp:575-589)
#    5  RBF     (Hayper_Euc_RBF, inverseWidth_RBF, Sigma_RBF; Kernel.cpp:414-431)
THETA0_EXP = np.array([0.5, 0.9, 0.2, 0.016])
THETA0_RBF = np.array([0.5, 0.9, 0.5, 0.2, 0.016])


def kernel_of(theta):
    return {10: "ExpAns", 4: "Exp", 5: "RBF"}[len(theta)]


def sigma_of(theta):
    """The kernel's amplitude parameter: Sigma_ExpAns / Sigma_Exp / Sigma_RBF (third from the end in every layout)."""
    return theta[6] if len(theta) == 10 else theta[-3]


# ----------------------------------------------------------------------------------
# A. symmetric standardisation  (Control.h:46-73, Control.cpp:299-324)
# ----------------------------------------------------------------------------------
def statistics_calc(X, y):
    """Control::StatisticsCalc (Control.h:46-73): row 0 = y, rows 1..D = X columns."""
    n, D = X.shape
    st = {
        "MaxTotalin": float(X.max()), "MinTotalin": float(X.min()),
        "MaxTotalo": float(y.max()), "MinTotalo": float(y.min()),
        "Min": np.zeros(D + 1), "Max": np.zeros(D + 1), "Mean": np.zeros(D + 1), "Std": np.zeros(D + 1),
    }
    cols = [y.reshape(-1)] + [X[:, j] for j in range(D)]
    for i, c in enumerate(cols):
        st["Min"][i] = c.min()
        st["Max"][i] = c.max()
        st["Mean"][i] = c.sum() / n
        st["Std"][i] = math.sqrt(((c - st["Mean"][i]) ** 2).sum() / (n - 1))
    return st


def prep_symmetric_params(st, D):
    """Control::prep_symmetric, train mode (Control.cpp:301-316): params[:,0]=centre, [:,1]=half-range.
    The first THREE X columns share one centre/half-range from the global X min/max."""
    params = np.zeros((D + 1, 2))
    params[0, 0] = 0.5 * (st["MaxTotalo"] + st["MinTotalo"])
    params[0, 1] = 0.5 * (st["MaxTotalo"] - st["MinTotalo"])
    for j in range(3):
        params[j + 1, 0] = 0.5 * (st["MaxTotalin"] + st["MinTotalin"])
        params[j + 1, 1] = 0.5 * (st["MaxTotalin"] - st["MinTotalin"])
    for j in range(3, D):
        params[j + 1, 0] = 0.5 * (st["Max"][j + 1] + st["Min"][j + 1])
        params[j + 1, 1] = 0.5 * (st["Max"][j + 1] - st["Min"][j + 1])
    return params


def apply_standardise(X, y, params):
    """Control.cpp:318-323."""
    Xs = np.empty_like(X)
    for j in range(X.shape[1]):
        Xs[:, j] = (X[:, j] - params[j + 1, 0]) / params[j + 1, 1]
    ys = (y - params[0, 0]) / params[0, 1]
    return Xs, ys


def standardise_train(X, y):
    st = statistics_calc(X, y)
    params = prep_symmetric_params(st, X.shape[1])
    Xs, ys = apply_standardise(X, y, params)
    return Xs, ys, params, st


def post_mean(mu, params):
    """Control::postData (Control.cpp:218): y = mu*half + centre."""
    return mu * params[0, 1] + params[0, 0]


def post_std(var, params):
    """Control::postData_var (Control.cpp:253-254): sqrt(var*half^2)."""
    return np.sqrt(var * params[0, 1] ** 2)


# ----------------------------------------------------------------------------------
# B. rotation / anisotropy matrix  (Kernel.cpp:1402-1425)
# ----------------------------------------------------------------------------------
def rot_matrix(alpha, beta, teta):
    """Rot (Kernel.cpp:1402-1410); libm sin/cos (math.*), the same routines the C++ host calls."""
    sa, ca = math.sin(alpha), math.cos(alpha)
    sb, cb = math.sin(beta), math.cos(beta)
    st, ct = math.sin(teta), math.cos(teta)
    R = np.zeros((3, 3))
    R[0, 0] = ca * ct + sa * sb * st
    R[0, 1] = -sa * ct + ca * sb * st
    R[0, 2] = -cb * st
    R[1, 0] = sa * cb
    R[1, 1] = ca * cb
    R[1, 2] = sb
    R[2, 0] = ca * st - sa * sb * ct
    R[2, 1] = -sa * st - ca * sb * ct
    R[2, 2] = cb * ct
    return R


def sig_inv(theta):
    """sigInv = Rot*lambda*Rot.t() (Kernel.cpp:1417-1425).  Evaluated in a DEFINED order shared with the
    product's host code (gp_ss_ak_b200/host/gpss_params.h): T = Rot*diag(l) element-wise, then
    S(i,j) = (T(i,0)*Rot(j,0) + T(i,1)*Rot(j,1)) + T(i,2)*Rot(j,2), every operation individually rounded."""
    R = rot_matrix(theta[0], theta[2], theta[4])
    lam = (theta[1], theta[3], theta[5])
    S = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            t0 = R[i, 0] * lam[0]
            t1 = R[i, 1] * lam[1]
            t2 = R[i, 2] * lam[2]
            S[i, j] = (t0 * R[j, 0] + t1 * R[j, 1]) + t2 * R[j, 2]
    return S


def sig_inv_full(theta, d):
    """The d x d sigInv of MahaDist: for the 4-column (rock-type) branch Rot(3,3) = 1 and lambda(3,3) = InversewidthR
    (Kernel.cpp:1411-1424), so the matrix is block diagonal: [sig_inv(theta), theta[7]]."""
    S3 = sig_inv(theta)
    if d == 3:
        return S3
    S = np.zeros((4, 4))
    S[:3, :3] = S3
    S[3, 3] = theta[7]
    return S


def seq_colsum(X):
    """Column sums accumulated strictly in row order (defined order shared with the product host code)."""
    s = np.zeros(X.shape[1])
    for j in range(X.shape[1]):
        acc = 0.0
        col = X[:, j]
        for v in col.tolist():
            acc += v
        s[j] = acc
    return s


def centre(X1, X2, sums1=None, sums2=None):
    """mX2 of MahaDist (Kernel.cpp:1391-1392): n/(n+m)*sum(X1)/n + m/(n+m)*sum(X2)/m."""
    n, m = X1.shape[0], X2.shape[0]
    s1 = seq_colsum(X1) if sums1 is None else sums1
    s2 = seq_colsum(X2) if sums2 is None else sums2
    mX1 = (float(n) / (n + m)) * s1 / n
    return (float(m) / (n + m)) * s2 / m + mX1


# ----------------------------------------------------------------------------------
# B/C/D. covariance, literal (BLAS) form  (Kernel.cpp:1370-1435, 856-882, 362-367, 140-154)
# ----------------------------------------------------------------------------------
def maha_dist_blas(X1, X2, theta):
    """MahaDist exactly as written: centred copies, X*sigInv via dgemm, expansion-form D2 via dgemm, clamp."""
    c = centre(X1, X2)
    S = sig_inv_full(theta, X1.shape[1])
    Z1 = (X1 - c) @ S
    Z2 = (X2 - c) @ S
    a1 = (Z1 * Z1)
    a2 = (Z2 * Z2)
    s1 = (a1[:, 0] + a1[:, 1]) + a1[:, 2]
    s2 = (a2[:, 0] + a2[:, 1]) + a2[:, 2]
    if X1.shape[1] == 4:
        s1 = s1 + a1[:, 3]
        s2 = s2 + a2[:, 3]
    a1, a2 = s1, s2
    D2 = (a1[:, None] + a2[None, :]) - 2.0 * (Z1 @ Z2.T)
    D2[D2 < 0] = 0.0
    return D2


def _fma_emul(a, b, c):
    """Exact fused multiply-add on float64 arrays via Dekker/Veltkamp error-free transformations
    (round(a*b+c) with one rounding).  Used when the C helper is unavailable."""
    # two-product (Veltkamp split)
    p = a * b
    SPL = 134217729.0
    ta = SPL * a
    ah = ta - (ta - a)
    al = a - ah
    tb = SPL * b
    bh = tb - (tb - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl  # a*b = p + e exactly
    # two-sum p + c
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)
    # s + (err + e): correct to 1 rounding except in rare double-rounding ties (not hit for these magnitudes;
    # the C helper oracle/c/oracle_kernels.c uses a hardware fma and is the authoritative path).
    return s + (err + e)


def transform_defined(X, c, S):
    """Z = (X - c) * S in the defined order: z_j = fma(d2, S[2,j], fma(d1, S[1,j], d0*S[0,j])), d = x - c."""
    d0 = X[:, 0] - c[0]
    d1 = X[:, 1] - c[1]
    d2 = X[:, 2] - c[2]
    Z = np.empty((X.shape[0], X.shape[1]))
    for j in range(3):
        Z[:, j] = _fma_emul(d2, S[2, j], _fma_emul(d1, S[1, j], d0 * S[0, j]))
    if X.shape[1] == 4:                      # 4-column branch: z_3 = (x_3 - c_3) * InversewidthR (block-diagonal sigInv)
        Z[:, 3] = (X[:, 3] - c[3]) * S[3, 3]
    return Z


def sqnorm_defined(Z):
    """a_i = fl(fl(z0^2 + z1^2) + z2^2) (sum(X%X,1), Kernel.cpp:1431)."""
    a = (Z[:, 0] * Z[:, 0] + Z[:, 1] * Z[:, 1]) + Z[:, 2] * Z[:, 2]
    if Z.shape[1] == 4:
        a = a + Z[:, 3] * Z[:, 3]
    return a


def maha_dist_defined(X1, X2, theta, c=None):
    """Defined-operation-order D2 (SURVEY.md section 7 hard part 1), the order the CUDA kernels use:
       c_ij = fma(z_i2, z_j2, fma(z_i1, z_j1, z_i0*z_j0));  D2 = max(0, fl(fl(a_i + a_j) - 2 c_ij))."""
    if c is None:
        c = centre(X1, X2)
    S = sig_inv_full(theta, X1.shape[1])
    Z1 = transform_defined(X1, c, S)
    Z2 = transform_defined(X2, c, S)
    a1 = sqnorm_defined(Z1)
    a2 = sqnorm_defined(Z2)
    cij = _fma_emul(Z1[:, 2][:, None], Z2[:, 2][None, :],
                    _fma_emul(Z1[:, 1][:, None], Z2[:, 1][None, :], Z1[:, 0][:, None] * Z2[:, 0][None, :]))
    if X1.shape[1] == 4:
        cij = _fma_emul(Z1[:, 3][:, None], Z2[:, 3][None, :], cij)
    D2 = (a1[:, None] + a2[None, :]) - 2.0 * cij
    D2[D2 < 0] = 0.0
    return D2


def eucl_dist_blas(X1, X2, hyp):
    """EuclDist as written (Kernel.cpp:1343-1368, mlA :1437-1441): centred copies, AX = X * exp(-2 log(hyp)),
    D2 = sum(AX1 % X1, 1) 1' + 1 sum(AX2 % X2, 1)' - 2 X1 AX2', negatives -> 0.  Net: |x - x'|^2 / hyp^2."""
    c = centre(X1, X2)
    Z1 = X1 - c
    Z2 = X2 - c
    f = math.exp(-2.0 * math.log(hyp))
    A1 = Z1 * f
    A2 = Z2 * f
    a1 = (A1 * Z1)
    a2 = (A2 * Z2)
    s1, s2 = a1[:, 0], a2[:, 0]
    for j in range(1, X1.shape[1]):
        s1 = s1 + a1[:, j]
        s2 = s2 + a2[:, j]
    D2 = (s1[:, None] + s2[None, :]) - 2.0 * (Z1 @ A2.T)
    D2[D2 < 0] = 0.0
    return D2


def eucl_dist_defined(X1, X2, hyp, c=None):
    """The defined-order form the CUDA kernels use for the isotropic kernels: the same pair-distance code as ExpAns with
    sigInv = (1/hyp) I, i.e. z = (x - c) * (1/hyp) and D2 = max(0, fl(fl(a_i + a_j) - 2 c_ij)) (maha_dist_defined)."""
    if c is None:
        c = centre(X1, X2)
    d = X1.shape[1]
    S = np.zeros((d, d))
    np.fill_diagonal(S, 1.0 / hyp)
    Z1 = transform_defined(X1, c, S)
    Z2 = transform_defined(X2, c, S)
    a1 = sqnorm_defined(Z1)
    a2 = sqnorm_defined(Z2)
    cij = _fma_emul(Z1[:, 2][:, None], Z2[:, 2][None, :],
                    _fma_emul(Z1[:, 1][:, None], Z2[:, 1][None, :], Z1[:, 0][:, None] * Z2[:, 0][None, :]))
    if d == 4:
        cij = _fma_emul(Z1[:, 3][:, None], Z2[:, 3][None, :], cij)
    D2 = (a1[:, None] + a2[None, :]) - 2.0 * cij
    D2[D2 < 0] = 0.0
    return D2


def compute_K(X1, X2, theta, dist="defined", c=None):
    """HybKerns::computeK = main kernel + Bias (Kernel.cpp:140-154, 362-367), bias added to EVERY element:
       ExpAns  K = Sigma^2 * exp(-sqrt(D2)) + Sigma_Bias, D2 = MahaDist            (Kernel.cpp:856-882)
       Exp     K = Sigma^2 * exp(-sqrt(D2)) + Sigma_Bias, D2 = EuclDist(hyp)       (Kernel.cpp:636-642)
       RBF     K = exp(-0.5 * inverseWidth * D2) * Sigma^2 + Sigma_Bias            (Kernel.cpp:482-488)"""
    kern = kernel_of(theta)
    if kern == "ExpAns":
        D2 = maha_dist_blas(X1, X2, theta) if dist == "blas" else maha_dist_defined(X1, X2, theta, c)
    else:
        D2 = eucl_dist_blas(X1, X2, theta[0]) if dist == "blas" else eucl_dist_defined(X1, X2, theta[0], c)
    sig = sigma_of(theta)
    var2 = sig * sig
    if kern == "RBF":
        K = np.exp(-0.5 * theta[1] * D2) * var2
    else:
        K = var2 * np.exp(-1 * np.sqrt(D2))
    K += theta[-2]
    return K, D2


# ----------------------------------------------------------------------------------
# F/G. Laplace/IRLS alpha, log-likelihood  (GP_Utils.cpp:180-416, 795-915, 1138-1162)
# ----------------------------------------------------------------------------------
def sign_ref(v):
    """ModelInf.h:14-20: sign(0) = -1."""
    return -1.0 if v <= 0 else 1.0


class OracleGP:
    """State-carrying restatement of GP_utils for Gaussian likelihood, zero mean (GP_Utils.cpp)."""

    def __init__(self, X, y, theta=None, dist="defined", literal=True, white=0.0, member2=None):
        # member2: a SECOND distance-based member of the Hyb sum (gp_ss_ak.cpp:146-175), as a parameter vector in the single-kernel
        # layout of its own kind (10 ExpAns / 4 Exp / 5 RBF) with the Sigma_Bias slot 0 and the sn2 slot ignored.  HybKerns::computeK adds
        # the members' K and D2 (Kernel.cpp:140-154); getGradients hands every member the SUMMED D2 (:156-169).
        self.member2 = None if member2 is None else np.array(member2, dtype=np.float64)
        # white: sum of the White members' Sigma_White (Kern_White, Kernel.cpp:180-270): K.diag() += white when computeK sees the same
        # point set twice (`X1(0) == X2(0) && X1.n_rows == X2.n_rows`, :261-262), prior variance += white (:222-225), gradient 0
        self.white = float(white)
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        self.n = self.X.shape[0]
        self.theta = np.array(THETA0 if theta is None else theta, dtype=np.float64)
        self.dist = dist
        self.literal = literal          # True: IRLS + Brent + 3 Cholesky (reference-literal); False: direct solve
        # Alpha starts at zero only because arma::Mat::resize zero-fills (GP_Utils.cpp:69)
        self.Alpha = np.zeros(self.n)
        self.K_ok = False
        self.alpha_ok = False
        self.chol_fail = False
        self.n_chol = 0
        self.n_psi = 0

    # -- parameter protocol (GP_Utils.cpp:101-157) --
    def get_pars(self):
        return self.theta.copy()

    def set_pars(self, th):
        self.K_ok = False           # setKUpdateStat(false) also clears AlphaUpStatus (GP_Utils.h:262-268)
        self.alpha_ok = False
        self.theta = np.array(th, dtype=np.float64).reshape(-1).copy()

    @property
    def sn2(self):
        return self.theta[-1]

    # -- kernel (GP_Utils.cpp:1100-1113) --
    def update_kernel(self):
        if not self.K_ok:
            self.K, self.D2 = compute_K(self.X, self.X, self.theta, self.dist)
            if self.member2 is not None:
                K2, D2b = compute_K(self.X, self.X, self.member2, self.dist)
                self.K = self.K + K2
                self.D2 = self.D2 + D2b
            if self.white != 0.0:
                self.K[np.diag_indices(self.n)] += self.white
            self.K_ok = True

    # -- likelihood terms (GP_Utils.cpp:398-416 and 795-839) --
    def _lik(self, f):
        sn2 = self.sn2
        ymmu = self.y - f
        self.lp = ymmu ** 2 * (-1 / (2 * sn2)) - math.log(2 * math.pi * sn2) / 2
        self.dlp = (1 / sn2) * ymmu
        self.d2lp = np.full(self.n, 1 / sn2)

    def _psi(self, alp):
        """PSI (GP_Utils.cpp:180-190); returns (psi, fval=K*alp)."""
        self.n_psi += 1
        f = self.K @ alp
        self._lik(f)
        return float(alp @ (0.5 * f) - self.lp.sum()), f

    def _chol_B(self):
        """chol(Sw Sw' % K + I) -> upper R (GP_Utils.cpp:876-888 / 898-910)."""
        self.n_chol += 1
        Sw = np.sqrt(self.d2lp)
        B = (Sw[:, None] * Sw[None, :]) * self.K + np.eye(self.n)
        try:
            R = sla.cholesky(B, lower=False, check_finite=False)
        except np.linalg.LinAlgError:
            self.chol_fail = True
            return None, Sw
        if not np.all(np.isfinite(np.diag(R))):
            self.chol_fail = True
            return None, Sw
        self.chol_fail = False
        return R, Sw

    @staticmethod
    def _solve_chol(R, b):
        """solve_chol (GP_Utils.cpp:841-845): solve(trimatl(R'), b) then solve(trimatu(R), .)."""
        t = sla.solve_triangular(R, b, trans="T", lower=False, check_finite=False)
        return sla.solve_triangular(R, t, trans="N", lower=False, check_finite=False)

    def _brentmin(self, dalpha):
        """brentmin (GP_Utils.cpp:229-381)."""
        Alpha = self.Alpha
        smin, smax, nmax, thr = 0.0, 2.0, 10, 1e-4
        counters = 0
        fa, _ = self._psi(Alpha + smin * dalpha); counters += 1
        fb, _ = self._psi(dalpha * smax + Alpha); counters += 1
        seps = math.sqrt(2.220446049250313e-16)
        c = 0.5 * (3.0 - math.sqrt(5.0))
        a, b = smin, smax
        v = a + c * (b - a)
        w = v
        xf = v
        d = 0.0
        e = 0.0
        x = xf
        fc, _ = self._psi(Alpha + x * dalpha); counters += 1
        fv = fc
        fw = fc
        xm = 0.5 * (a + b)
        tol1 = seps * abs(xf) + thr / 3.0
        tol2 = 2.0 * tol1
        while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
            gs = 1
            if abs(e) > tol1:
                gs = 0
                r = (xf - w) * (fc - fv)
                q = (xf - v) * (fc - fw)
                p = (xf - v) * q - (xf - w) * r
                q = 2.0 * (q - r)
                if q > 0.0:
                    p = -p
                q = abs(q)
                r = e
                e = d
                if (abs(p) < abs(0.5 * q * r)) and (p > q * (a - xf)) and (p < q * (b - xf)):
                    d = p / q
                    x = xf + d
                    if ((x - a) < tol2) or ((b - x) < tol2):
                        si = sign_ref(xm - xf) + (1.0 if (xm - xf) == 0 else 0.0)
                        d = tol1 * si
                else:
                    gs = 1
            if gs == 1:
                if xf >= xm:
                    e = a - xf
                else:
                    e = b - xf
                d = c * e
            si = sign_ref(d) + (1.0 if d == 0 else 0.0)
            sd = tol1 if abs(d) < tol1 else abs(d)
            x = xf + si * sd
            fu, _ = self._psi(dalpha * x + Alpha); counters += 1
            if fu <= fc:
                if x >= xf:
                    a = xf
                else:
                    b = xf
                v = w; fv = fw
                w = xf; fw = fc
                xf = x
                fc = fu
            else:
                if x < xf:
                    a = x
                else:
                    b = x
                if (fu <= fw) or (w == xf):
                    v = w; fv = fw
                    w = x; fw = fu
                elif (fu <= fv) or (v == xf) or (v == w):
                    v = x; fv = fu
            xm = 0.5 * (a + b)
            tol1 = seps * abs(xf) + thr / 3.0
            tol2 = 2.0 * tol1
            if counters >= nmax:
                break
        if (fa < fc) and (fa <= fb):
            xf = smin; fc = fa
        elif fb < fc:
            xf = smax; fc = fb
        fmin = fc
        Xc = dalpha * xf + Alpha
        self.Alpha = Xc
        _, Fv = self._psi(Xc)
        self.last_step = xf
        return fmin, Fv

    def _irls(self):
        """irls (GP_Utils.cpp:191-228)."""
        maxit, tol = 20, 1e-6
        psi_new, Fv = self._psi(self.Alpha)
        psi_old = math.inf
        it = 0
        self.irls_steps = []
        while (psi_old - psi_new) > tol and it < maxit:
            psi_old = psi_new
            it += 1
            B = Fv * self.d2lp + self.dlp
            r = self.K @ B
            R, Sw = self._chol_B()
            if self.chol_fail:
                return
            self.Lchol = R
            dalpha = self._solve_chol(R, Sw * r) * Sw
            dalpha = -1 * dalpha - self.Alpha + B
            psi_new, Fv = self._brentmin(dalpha)
            self.irls_steps.append(self.last_step)
        self.irls_its = it

    def update_alpha(self):
        if not self.alpha_ok:
            self.update_kernel()
            if self.literal:
                self._irls()
                if self.chol_fail:
                    return
            else:
                # the fixed point the IRLS converges to: alpha = (K + sn2 I)^-1 y
                self._lik(np.zeros(self.n))
                R, Sw = self._chol_B()
                if self.chol_fail:
                    return
                self.Lchol = R
                self.Alpha = self._solve_chol(R, self.y / self.sn2)
            self.alpha_ok = True

    def log_likelihood(self):
        """logLikelihood (GP_Utils.cpp:1138-1162): the NEGATIVE log marginal likelihood."""
        self.update_kernel()
        self.update_alpha()
        if self.chol_fail:
            return math.nan
        self.yhat = self.K @ self.Alpha
        self._lik(self.yhat)
        ydif = 0.5 * self.yhat
        if self.literal:
            R, Sw = self._chol_B()          # the reference factorises the same matrix again (GP_Utils.cpp:894-915)
            if self.chol_fail:
                return math.nan
            self.Lchol = R
        self.Sw = np.sqrt(self.d2lp)
        self.Lchol_db2 = float(np.log(np.diag(self.Lchol)).sum())
        self.L = float(self.Alpha @ ydif - self.lp.sum() + self.Lchol_db2)
        return self.L

    # -- H/I/J. gradient (GP_Utils.cpp:1164-1284, Kernel.cpp:886-1263, 370-377) --
    def grad_ll(self):
        L = self.log_likelihood()
        if self.chol_fail:
            return math.nan, np.full(len(self.theta), math.nan)
        n = self.n
        Sw = self.Sw
        R = self.Lchol
        # Q = B^-1 through two triangular solves with an n x n right-hand side (GP_Utils.cpp:1202-1205)
        Q = self._solve_chol(R, np.diag(Sw))
        Q = Q * ((1 / Sw)[:, None] * np.ones((1, n)))
        dW = 0.5 * (Q * self.K).sum(axis=1)
        d3lp = np.zeros(n)
        dfhat = dW * d3lp
        kmvm = (self.K @ dfhat) * Sw
        dahat = self._solve_chol(R, kmvm)
        dahat = -1 * dahat * Sw + dfhat
        # dhyp (GP_Utils.cpp:1164-1169)
        QW = Q * (self.d2lp[:, None] * np.ones((1, n))) - np.outer(self.Alpha, self.Alpha) \
            + np.outer(self.dlp, dahat) * 2.0
        self.QW = QW
        npar = len(self.theta)
        g = np.zeros(npar)
        kern = kernel_of(self.theta)
        if kern == "ExpAns":
            g[0:8] = expans_gradients_literal(self.X, self.theta, QW, self.dist)
        elif kern == "Exp":
            g[0:2] = exp_gradients_literal(self.theta, QW, self.D2)      # HybKerns hands the members the SUMMED D2 (Kernel.cpp:156-169)
        else:
            g[0:3] = rbf_gradients_literal(self.theta, QW, self.D2)
        g[npar - 2] = bias_gradient_literal(QW)
        if self.member2 is not None:
            k2 = kernel_of(self.member2)
            if k2 == "ExpAns":
                self.g2 = expans_gradients_literal(self.X, self.member2, QW, self.dist)
            elif k2 == "Exp":
                self.g2 = exp_gradients_literal(self.member2, QW, self.D2)
            else:
                self.g2 = rbf_gradients_literal(self.member2, QW, self.D2)
        # likelihood hyper-parameter (GP_Utils.cpp:1222-1235, 846-871)
        sn2 = self.sn2
        ymmu = self.y - self.yhat
        lp_dhyp = (1 / sn2) * ymmu ** 2 - 1
        dlp_dhyp = (-2 / sn2) * ymmu
        d2lp_dhyp = np.full(n, 2 / sn2)
        g_tmp0 = -1 * (dW @ d2lp_dhyp) - lp_dhyp.sum()
        B = self.K @ dlp_dhyp
        B0 = B.copy()
        B = self._solve_chol(R, B * Sw) * Sw
        B = -1 * B + B0
        g_tmp0 += -1.0 * (dfhat @ B)
        g[npar - 1] = g_tmp0
        self.dW = dW
        return L, g

    # -- K. prediction (GP_Utils.cpp:943-1041) --
    def predict(self, Xs):
        Xs = np.ascontiguousarray(Xs, dtype=np.float64)
        kX, _ = compute_K(self.X, Xs, self.theta, self.dist)      # n x m, centre uses BOTH sets (Kernel.cpp:1391)
        if self.member2 is not None:
            kX = kX + compute_K(self.X, Xs, self.member2, self.dist)[0]
        if self.white != 0.0 and Xs.shape[0] == self.n and Xs[0, 0] == self.X[0, 0]:      # Kern_White::computeK(X_train, X_test), Kernel.cpp:261-262
            kX[np.diag_indices(self.n)] += self.white
        self.update_alpha()
        mu = kX.T @ self.Alpha
        kD = np.full(Xs.shape[0], sigma_of(self.theta) ** 2 + self.theta[-2] + self.white)      # diag_Compute (Kernel.cpp:782, 331, 449, 594, 222-225)
        if self.member2 is not None:
            kD = kD + sigma_of(self.member2) ** 2
        self.log_likelihood()                                         # GP_Utils.cpp:980
        Wh = np.sqrt(self.d2lp)
        LKs = kX * Wh[:, None]
        LKs = self._solve_chol(self.Lchol, LKs)
        LKs = LKs * Wh[:, None]
        LKs *= kX
        var = kD - LKs.sum(axis=0)
        self.var_raw = var.copy()
        var = var_postprocess(var, self.sn2)
        return mu, var


def var_postprocess(var_raw, sn2):
    """GP_Utils.cpp:1001-1003 then 1033-1040, literally:
           uvec ind = varSigma < 0;                          // 0/1 FLAGS ...
           varSigma.elem(ind) = zeros<mat>(ind.n_rows, ind.n_cols);   // ... used as INDICES
       so element 0 is zeroed when any entry is non-negative, element 1 when any entry is negative, negative entries
       are not clamped; then `varSigma += sn2` unless sn2 == 1.0.  (The compiled reference confirms it: var[0] == sn2
       in every tests/golden/ref_*.npz record.)"""
    var = np.array(var_raw, dtype=np.float64).reshape(-1).copy()
    ind = (var < 0).astype(np.int64)
    if ind.max(initial=0) >= var.shape[0]:
        raise IndexError("Mat::elem(): index out of bounds")      # the reference aborts here
    var[ind] = 0.0
    if sn2 != 1.0:
        var = var + sn2
    return var


def s_matrices(theta):
    """S and the six S_p of Kern_ExpAnisotropic::getGradients, entry by entry (Kernel.cpp:946-1166),
    including the quirk that S_angle(0,0) carries no factor 2 on its z term (Kernel.cpp:1003-1011)."""
    al, be, te = theta[0], theta[2], theta[4]
    l = (theta[1], theta[3], theta[5])
    sa, ca, sb, cb, st, ct = math.sin(al), math.cos(al), math.sin(be), math.cos(be), math.sin(te), math.cos(te)
    Rot = rot_matrix(al, be, te)
    Ra = np.zeros((3, 3)); Rb = np.zeros((3, 3)); Rt = np.zeros((3, 3))
    Ra[0, 0] = -sa * ct + ca * sb * st
    Rb[0, 0] = sa * cb * st
    Rt[0, 0] = -ca * st + sa * sb * ct
    Ra[0, 1] = -ca * ct - sa * sb * st
    Rb[0, 1] = ca * cb * st
    Rt[0, 1] = sa * st + ca * sb * ct
    Ra[0, 2] = 0.0
    Rb[0, 2] = sb * st
    Rt[0, 2] = -cb * ct
    Ra[1, 0] = ca * cb
    Rb[1, 0] = -sa * sb
    Rt[1, 0] = 0.0
    Ra[1, 1] = -sa * cb
    Rb[1, 1] = -ca * sb
    Rt[1, 1] = 0.0
    Ra[1, 2] = 0.0
    Rb[1, 2] = cb
    Rt[1, 2] = 0.0
    Ra[2, 0] = -sa * st - ca * sb * ct
    Rb[2, 0] = -sa * cb * ct
    Rt[2, 0] = ca * ct + sa * sb * st
    Ra[2, 1] = -ca * st + sa * sb * ct
    Rb[2, 1] = -ca * cb * ct
    Rt[2, 1] = -sa * ct + ca * sb * st
    Ra[2, 2] = 0.0
    Rb[2, 2] = -sb * ct
    Rt[2, 2] = -cb * st

    S = np.zeros((3, 3))
    Sd = {k: np.zeros((3, 3)) for k in ("a", "b", "t")}
    SL = [np.zeros((3, 3)) for _ in range(3)]
    dR = {"a": Ra, "b": Rb, "t": Rt}
    # (0,0): the quirk
    S[0, 0] = l[0] * Rot[0, 0] ** 2 + l[1] * Rot[0, 1] ** 2 + l[2] * Rot[0, 2] ** 2
    for k, D in dR.items():
        Sd[k][0, 0] = l[0] * 2 * Rot[0, 0] * D[0, 0] + l[1] * 2 * Rot[0, 1] * D[0, 1] + l[2] * Rot[0, 2] * D[0, 2]
    for q in range(3):
        SL[q][0, 0] = Rot[0, q] ** 2
    # remaining upper-triangle entries: the regular product rule
    for (i, j) in ((0, 1), (0, 2), (1, 1), (1, 2), (2, 2)):
        S[i, j] = l[0] * Rot[i, 0] * Rot[j, 0] + l[1] * Rot[i, 1] * Rot[j, 1] + l[2] * Rot[i, 2] * Rot[j, 2]
        for k, D in dR.items():
            Sd[k][i, j] = (l[0] * D[i, 0] * Rot[j, 0] + l[0] * Rot[i, 0] * D[j, 0]
                           + l[1] * D[i, 1] * Rot[j, 1] + l[1] * Rot[i, 1] * D[j, 1]
                           + l[2] * D[i, 2] * Rot[j, 2] + l[2] * Rot[i, 2] * D[j, 2])
        for q in range(3):
            SL[q][i, j] = Rot[i, q] * Rot[j, q]
    for (i, j) in ((1, 0), (2, 0), (2, 1)):
        S[i, j] = S[j, i]
        for k in Sd:
            Sd[k][i, j] = Sd[k][j, i]
        for q in range(3):
            SL[q][i, j] = SL[q][j, i]
    # parameter order: AngleX, iWx, AngleY, iWy, AngleZ, iWz  (Kernel.cpp:1192-1233)
    return S, [Sd["a"], SL[0], Sd["b"], SL[1], Sd["t"], SL[2]]


def expans_gradients_literal(X, theta, QW, dist="defined"):
    """Kern_ExpAnisotropic::getGradients in its matrix form (Kernel.cpp:886-1263), 3-D branch."""
    n = X.shape[0]
    var2 = theta[6] * theta[6]
    if dist == "blas":
        DD2 = maha_dist_blas(X, X, theta)
    else:
        DD2 = maha_dist_defined(X, X, theta)
    S, Sp = s_matrices(theta)
    Qs = var2 * QW
    SD2 = np.sqrt(DD2)
    KD2 = np.exp(-1 * SD2)
    with np.errstate(divide="ignore", invalid="ignore"):
        dk_tmp = -0.5 / SD2
    dk_tmp[SD2 == 0] = 0
    dk = np.exp(-1.0 * SD2) * dk_tmp
    np.fill_diagonal(dk, 0.0)
    Rm = Qs * dk
    g = np.zeros(8)
    d = X.shape[1]
    XX = X * X
    for p in range(6):
        M = np.zeros((d, d))
        M[:3, :3] = S * Sp[p]                           # Hadamard product S % S_p (Kernel.cpp:1192); 4th row/column of S_p are zero
        rowq = (2.0 * XX @ M).sum(axis=1)
        Di2 = rowq[:, None] + rowq[None, :] - 4.0 * ((X @ M) @ X.T)
        g[p] = float((Rm * Di2).sum())
    g[6] = 2.0 * float((KD2 * QW).sum()) * theta[6]     # Kernel.cpp:1239-1242
    if d == 4:
        # S(3,3) = 1 and InversewidthRock(3,3) = 1 (Kernel.cpp:1169-1173), so S % InversewidthRock = e4 e4':
        # Di2 = 2 x_i4^2 + 2 x_j4^2 - 4 x_i4 x_j4.  [quirk] RColon was reassigned to KD2 = exp(-s) for the Sigma gradient
        # just above (Kernel.cpp:1240) and is NOT restored: g[7] = -2 sum(exp(-s) % Di2) / n -- QW does not enter at all
        # (Kernel.cpp:1246-1255; confirmed by the compiled reference, tests/golden/ref_rock_n300.npz).
        M = np.zeros((4, 4))
        M[3, 3] = 1.0
        rowq = (2.0 * XX @ M).sum(axis=1)
        Di2 = rowq[:, None] + rowq[None, :] - 4.0 * ((X @ M) @ X.T)
        g[7] = -2.0 * float((KD2 * Di2).sum()) / n
    else:
        g[7] = 0.0                                      # Kernel.cpp:1256-1257
    return g


def exp_gradients_literal(theta, QW, D2):
    """Kern_Exponential::getGradients (Kernel.cpp:646-695) on the D2 it is handed:
       dk = exp(-s) % (-0.5 / s) with the DIAGONAL zeroed only -- an off-diagonal s == 0 (duplicate points) gives
       -inf * 0 = NaN, as in the reference;  g[0] = sum(var2 QW % dk % D2);  g[1] = Sigma * sum(exp(-s) % (QW % exp(-s)))
       [quirk: exp(-s) enters squared]."""
    sig = theta[1]
    var2 = sig * sig
    SD2 = np.sqrt(D2)
    KD2 = np.exp(-1 * SD2)
    with np.errstate(divide="ignore", invalid="ignore"):
        dk = np.exp(-1 * SD2) * (-0.5 / SD2)
    np.fill_diagonal(dk, 0.0)
    Rm = (var2 * QW) * dk
    with np.errstate(invalid="ignore"):
        g0 = float((Rm * D2).sum())
    Q = QW * KD2
    g1 = float((KD2 * Q).sum()) * sig
    return np.array([g0, g1])


def rbf_gradients_literal(theta, QW, D2):
    """Kern_RBF::getGradients (Kernel.cpp:491-541): with KD2 = exp(-0.5 w D2), Qs = Sigma^2 QW,
       g[0] = (-2 sum(Qs % (KD2 * (-w/2)) % D2)) / 2;  g[1] = sum(-0.5 (Qs % KD2) % D2) / 2;
       g[2] = ((sum(QW % KD2) Sigma + sum((QW % KD2)') Sigma) * Sigma) / 2."""
    w, sig = theta[1], theta[2]
    var2 = sig * sig
    Qs = var2 * QW
    KD2 = np.exp(-0.5 * w * D2)
    dk = np.exp(-w / 2 * D2) * (-w / 2)
    g1 = -2.0 * float(((Qs * dk) * D2).sum())
    g2 = float((-0.5 * (Qs * KD2) * D2).sum())
    Q = QW * KD2
    g3 = (float((Q * sig).sum()) + float((Q.T * sig).sum())) * sig
    return np.array([g1 / 2, g2 / 2, g3 / 2])


def bias_gradient_literal(QW):
    """Kern_Bias::getGradients (Kernel.cpp:370-377): vec(QW) . vec(eye) = trace(QW)."""
    return float(np.trace(QW))


def expans_gradients_fused(X, theta, QW, DD2):
    """Algebraically identical one-pass form used by the CUDA kernel:
       g[p] = 4 sum_i q_p(x_i) omega_i - 4 <M_p, T>,  omega = w 1,  T = X' w X,  q_p(x) = sum_k x_k^2 rho_pk."""
    var2 = theta[6] * theta[6]
    S, Sp = s_matrices(theta)
    SD2 = np.sqrt(DD2)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = var2 * QW * np.exp(-SD2) * (-0.5 / SD2)
    w[SD2 == 0] = 0
    np.fill_diagonal(w, 0.0)
    omega_r = w.sum(axis=1)
    omega_c = w.sum(axis=0)
    X3 = X[:, :3]
    T = X3.T @ w @ X3
    g = np.zeros(8)
    XX = X3 * X3
    for p in range(6):
        M = S * Sp[p]
        rho = M.sum(axis=0)        # (XX @ M).sum(axis=1) = XX @ (M 1)
        q = XX @ M.sum(axis=1)
        g[p] = 2.0 * float(q @ omega_r) + 2.0 * float(q @ omega_c) - 4.0 * float((M * T).sum())
    g[6] = 2.0 * theta[6] * float((QW * np.exp(-SD2)).sum())
    if X.shape[1] == 4:            # g[7] = -4/n sum_ij exp(-s_ij) (x_i4 - x_j4)^2   (no QW: see expans_gradients_literal)
        dx = X[:, 3][:, None] - X[:, 3][None, :]
        g[7] = -4.0 * float((np.exp(-SD2) * dx * dx).sum()) / X.shape[0]
    return g


# ----------------------------------------------------------------------------------
# convenience entry points used by tests / bench
# ----------------------------------------------------------------------------------
def nlml_and_grad(X, y, theta, dist="defined", literal=True):
    gp = OracleGP(X, y, theta, dist=dist, literal=literal)
    L, g = gp.grad_ll()
    return L, g, gp


def nlml_direct(X, y, theta, dist="defined"):
    """Textbook value the reference's objective equals at the IRLS fixed point:
       0.5 y' alpha + 0.5 log det(K + sn2 I) + n/2 log(2 pi)."""
    K, _ = compute_K(X, X, theta, dist)
    n = X.shape[0]
    sn2 = theta[9]
    Lc = sla.cholesky(K + sn2 * np.eye(n), lower=True, check_finite=False)
    alpha = sla.cho_solve((Lc, True), y, check_finite=False)
    return 0.5 * float(y @ alpha) + float(np.log(np.diag(Lc)).sum()) + 0.5 * n * math.log(2 * math.pi), alpha


def white_fixture_cases(z):
    """(tag, k, 10-slot theta of the C ABI, white) for the parameter vectors of tests/golden/ref_white_n300.npz:
    Hyb{White, Bias} = [Sigma_White, Sigma_Bias, sn2] (no distance member: the ExpAns slots run with Sigma = 0) and
    Hyb{ExpAns, White, Bias} = [ExpAns x 8, Sigma_White, Sigma_Bias, sn2]."""
    out = []
    for tag in ("w", "ew"):
        for k in range(2):
            th = z["%s_foreign_theta_%d" % (tag, k)].ravel()
            if tag == "w":
                t10 = THETA0.copy()
                t10[6], t10[8], t10[9] = 0.0, th[1], th[2]
                out.append((tag, k, t10, float(th[0])))
            else:
                out.append((tag, k, np.concatenate([th[:8], th[9:]]), float(th[8])))
    return out


NPAR_MEMBER = {"ExpAns": 8, "Exp": 2, "RBF": 3}
KIND_CODE = {"ExpAns": 0, "Exp": 1, "RBF": 2}


def split_sum2_theta(combo, th):
    """Parameter vector of Hyb{a, b, Bias} in the reference's order [a's parameters, b's parameters, Sigma_Bias, sn2] ->
    (theta of member a in its single-kernel layout [.., Sigma_Bias, sn2], member b in its layout with the bias slot 0, names)."""
    a, b = combo.split("+")
    na, nb = NPAR_MEMBER[a], NPAR_MEMBER[b]
    th = np.asarray(th, dtype=np.float64).ravel()
    t1 = np.concatenate([th[:na], th[na + nb:]])
    t2 = np.concatenate([th[na:na + nb], [0.0, th[-1]]])
    return t1, t2, a, b
