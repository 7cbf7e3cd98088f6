"""Fixtures for the White member of the additive covariance (Kern_White, /root/reference/Kernel.cpp:180-270) from the UNMODIFIED
reference (oracle/_ref/ref_driver).  Build container only; the .npz is committed.

    python tests/golden/make_ref_white.py

The reference cannot differentiate a Hyb kernel that holds a White member -- Kernels::getGradients' default calls itself
(Kernel.h:56-59) and the process dies of stack overflow -- so the dumps are taken with GPSS_REF_NOGRAD=1: objective, Alpha, diag(K),
K columns and predictions, for Hyb{White, Bias} and Hyb{ExpAns, White, Bias}, each on a foreign test set AND on the training set
itself (where Kern_White::computeK's `X1(0) == X2(0) && equal rows` test puts Sigma_White on the cross-covariance diagonal).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gp_ss_ak_b200 import datagen          # noqa: E402
from make_ref_golden import DRIVER, THETA0, parse_dump          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(kernel, thetas, X, y, Xt, yt):
    with tempfile.TemporaryDirectory() as d:
        datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
        datagen.write_data_file(os.path.join(d, "test.txt"), Xt, yt)
        with open(os.path.join(d, "thetas.txt"), "w") as f:
            for th in thetas:
                f.write(" ".join("%.17g" % v for v in th) + "\n")
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1", GPSS_REF_KERNEL=kernel, GPSS_REF_NOGRAD="1")
        out = subprocess.run([DRIVER, os.path.join(d, "train.txt"), os.path.join(d, "test.txt"), os.path.join(d, "thetas.txt"), "0", d],
                             capture_output=True, text=True, check=True, env=env)
        return parse_dump(out.stdout)


if __name__ == "__main__":
    n, seed = 300, 5
    X, y = datagen.drillholes(n, seed)
    Xt, _ = datagen.drillholes(40, seed + 50)
    Xt = np.concatenate([Xt, X[:10]])
    yt = np.concatenate([datagen.grade_at(Xt[:40], seed), y[:10]])
    rng = np.random.default_rng(11)
    rec = {"X_raw": X, "y_raw": y, "Xt_raw": Xt, "yt_raw": yt}
    # parameter order = member order: [Sigma_White, Sigma_Bias, sn2] and [ExpAns x 8, Sigma_White, Sigma_Bias, sn2]
    th_w = [np.array([0.10, 0.2, 0.016]), np.array([0.23, 0.15, 0.05])]
    th_ew0 = np.concatenate([THETA0[:8], [0.10], THETA0[8:]])
    th_ew = [th_ew0, np.clip(th_ew0 * rng.uniform(0.8, 1.25, 11), 1e-4, 6.0)]
    for tag, kernel, ths in (("w", "White", th_w), ("ew", "ExpAns+White", th_ew)):
        for tset, (A, b) in (("foreign", (Xt, yt)), ("self", (X, y))):
            r = run(kernel, ths, X, y, A, b)
            for k in range(len(ths)):
                for key in ("theta", "nlml", "alpha", "K_diag", "K_col0", "K_col17", "mu", "var", "yhat_raw", "std_raw"):
                    rec["%s_%s_%s_%d" % (tag, tset, key, k)] = r["%s_%d" % (key, k)]
            rec["Xs"] = r["Xs"]; rec["ys"] = r["ys"]; rec["params"] = r["params"]
            if tset == "foreign":
                rec["Xt"] = r["Xt"]
            print(tag, tset, "nlml", [r["nlml_%d" % k] for k in range(len(ths))], "var[:3]", r["var_0"][:3].ravel())
    # the reference's own command line: it prints the initial model and its objective, then dies in the first gradient (rc 139)
    cli = os.path.join(ROOT, "oracle", "_ref", "gp_ss_ak")
    with tempfile.TemporaryDirectory() as d:
        datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
        for tag, ks in (("w", ["-k", "White"]), ("ew", ["-k", "ExpAns", "-k", "White"])):
            tr = subprocess.run([cli, "-v", "3", "-pm", "1", "train"] + ks + ["-kn", "1", "-o", "LBFGS", "-#", "2", os.path.join(d, "train.txt"),
                                 os.path.join(d, "m_" + tag)], capture_output=True, text=True, stdin=subprocess.DEVNULL, cwd=d,
                                env=dict(os.environ, OPENBLAS_NUM_THREADS="1", MALLOC_PERTURB_="255"))
            rec["cli_%s_stdout" % tag] = np.array(tr.stdout)
            rec["cli_%s_rc" % tag] = tr.returncode
            print("cli", tag, "rc", tr.returncode, [l for l in tr.stdout.splitlines() if "Log likelihood" in l])
        rec["train_file_text"] = np.array(open(os.path.join(d, "train.txt")).read())
    np.savez_compressed(os.path.join(HERE, "ref_white_n300.npz"), **rec)
