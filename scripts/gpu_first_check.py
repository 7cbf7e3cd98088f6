"""First on-GPU check: kernel-level and end-to-end parity against the oracle (scratch harness, not a test)."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen
from oracle import gpss_oracle as O
import scipy.linalg as sla

rng = np.random.default_rng(0)
out = {}
# --- GEMM ---
for tile in (0, 1):
    M, N, K = 384, 256, 160
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K)); C0 = rng.standard_normal((M, N))
    C, ms = G.test_gemm_nt(A, B, tile=tile)
    err = np.abs(C - A @ B.T).max()
    C2, ms = G.test_gemm_nt(A, B, C=C0, tile=tile)
    err2 = np.abs(C2 - (C0 - A @ B.T)).max()
    print("gemm tile", tile, "err", err, err2, flush=True)
    out["gemm_err_%d" % tile] = [float(err), float(err2)]
for (M, N, K) in ((8192, 8192, 512), (8192, 8192, 4096), (16384, 16384, 512)):
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K))
    for rep in range(2):
        C, ms = G.test_gemm_nt(A, B, tile=0)
    tf = 2.0 * M * N * K / ms * 1e-9
    print("gemm perf", M, N, K, "ms", ms, "TFLOP/s", tf, flush=True)
    out["gemm_perf_%d_%d_%d" % (M, N, K)] = [ms, tf]
    if K == 512 and M == 8192:
        err = np.abs(C[:256, :256] - A[:256] @ B[:256].T).max()
        print("  spot err", err)
# --- potrf ---
for n in (300, 1024, 4096):
    Am = rng.standard_normal((n, n)); S = Am @ Am.T + n * np.eye(n)
    L, ld, ms, rc = G.test_potrf(S)
    Lr = sla.cholesky(S, lower=True)
    print("potrf n", n, "rc", rc, "err", np.abs(L - Lr).max() / np.abs(Lr).max(), "logdet", ld, np.log(np.diag(Lr)).sum(), "ms", ms, flush=True)
    out["potrf_err_%d" % n] = float(np.abs(L - Lr).max() / np.abs(Lr).max())
# --- end to end vs oracle ---
for n in (500, 2000):
    X, y = datagen.drillholes(n, 0)
    Xs, ys, params = datagen.standardise_symmetric(X, y)
    th = O.THETA0.copy()
    Kd, D2d = G.compute_K(th, Xs, Xs)
    Ko, D2o = O.compute_K(Xs, Xs, th)
    print("n", n, "D2 bit-equal frac", (D2d == D2o).mean(), "max abs diff", np.abs(D2d - D2o).max(), "K maxdiff", np.abs(Kd - Ko).max(), flush=True)
    t0 = time.time()
    Lo, go, gp = O.nlml_and_grad(Xs, ys, th, literal=True)
    t_or = time.time() - t0
    m = G.GpssModel(Xs, ys)
    m.set_theta(th)
    Lg = m.nlml()
    print("  nlml gpu", Lg, "oracle", Lo, "rel", abs(Lg - Lo) / abs(Lo), flush=True)
    Lg2, gg = m.nlml_grad()
    rel = np.abs(gg - go) / np.maximum(np.abs(go), 1e-300)
    print("  grad gpu", gg)
    print("  grad ora", go)
    print("  grad rel", rel, flush=True)
    al = m.alpha()
    print("  alpha rel", np.linalg.norm(al - gp.Alpha) / np.linalg.norm(gp.Alpha))
    Xt, yt = datagen.drillholes(300, 7)
    Xt = np.concatenate([Xt, X[:50]], axis=0)
    Xts = (Xt - params[1:, 0]) / params[1:, 1]
    mu_o, var_o = gp.predict(Xts)
    mu_g, var_g = m.predict(Xts)
    print("  pred mu maxabs", np.abs(mu_g - mu_o).max(), "var maxabs", np.abs(var_g - var_o).max(), "oracle time", t_or, flush=True)
    out["e2e_%d" % n] = dict(nlml_rel=float(abs(Lg - Lo) / abs(Lo)), grad_rel=[float(v) for v in rel],
                            mu=float(np.abs(mu_g - mu_o).max()), var=float(np.abs(var_g - var_o).max()))
    m.close()
# --- timing at larger n ---
for n in (8192, 20000):
    X, y = datagen.drillholes(n, 1)
    Xs, ys, params = datagen.standardise_symmetric(X, y)
    m = G.GpssModel(Xs, ys)
    m.set_profiling(True)
    th = O.THETA0.copy()
    for rep in range(2):
        th2 = th * (1 + 0.01 * rep)
        m.set_theta(th2)
        t0 = time.time(); L, g = m.nlml_grad(); dt = time.time() - t0
        ph = m.phase_ms()
        print("n", n, "rep", rep, "nlml", L, "wall s", dt, "phases ms", [round(v, 2) for v in ph[:8]], flush=True)
    npad = m.padded_n()
    out["time_%d" % n] = dict(wall=dt, phases=[float(v) for v in ph[:8]],
                             potrf_tflops=npad ** 3 / 3 / (ph[1] * 1e-3) * 1e-12,
                             trtri_tflops=npad ** 3 / 3 / (ph[3] * 1e-3) * 1e-12,
                             lauum_tflops=npad ** 3 / 3 / (ph[4] * 1e-3) * 1e-12)
    print("  TFLOP/s potrf/trtri/lauum", out["time_%d" % n]["potrf_tflops"], out["time_%d" % n]["trtri_tflops"], out["time_%d" % n]["lauum_tflops"])
    m.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/first_check.json", "w"), indent=1)
print("DONE")
