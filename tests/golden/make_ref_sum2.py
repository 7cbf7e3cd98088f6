"""Fixtures for Hyb sums of TWO distance-based members (gp_ss_ak train -k ExpAns -k RBF ..., /root/reference/gp_ss_ak.cpp:146-175;
HybKerns::computeK / getGradients, Kernel.cpp:140-169) from the UNMODIFIED reference (oracle/_ref/ref_driver with
GPSS_REF_KERNEL=ExpAns+RBF / Exp+ExpAns / RBF+Exp).  Build container only; the .npz is committed.

    python tests/golden/make_ref_sum2.py
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
P_EXPANS = THETA0[:8]
P_EXP = np.array([0.5, 0.9])               # Hayper_Euc_Exp, Sigma_Exp (Kernel.cpp:585-589)
P_RBF = np.array([0.5, 0.9, 0.5])          # Hayper_Euc_RBF, inverseWidth_RBF, Sigma_RBF (Kernel.cpp:425-431)
PARS = {"ExpAns": P_EXPANS, "Exp": P_EXP, "RBF": P_RBF}

if __name__ == "__main__":
    n, seed = 300, 8
    X, y = datagen.drillholes(n, seed)
    Xt, _ = datagen.drillholes(40, seed + 50)
    Xt = np.concatenate([Xt, X[:10]])
    yt = np.concatenate([datagen.grade_at(Xt[:40], seed), y[:10]])
    rng = np.random.default_rng(23)
    rec = {"X_raw": X, "y_raw": y, "Xt_raw": Xt, "yt_raw": yt}
    combos = [("ExpAns", "RBF"), ("Exp", "ExpAns"), ("RBF", "Exp")]
    rec["combos"] = np.array(["+".join(c) for c in combos])
    for a, b in combos:
        th0 = np.concatenate([PARS[a], PARS[b], [0.2, 0.016]])          # member a, member b, Sigma_Bias, sn2
        ths = [th0, np.clip(th0 * rng.uniform(0.8, 1.25, th0.shape[0]), 1e-4, 6.0)]
        with tempfile.TemporaryDirectory() as d:
            datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
            datagen.write_data_file(os.path.join(d, "test.txt"), Xt, yt)
            with open(os.path.join(d, "thetas.txt"), "w") as f:
                for th in ths:
                    f.write(" ".join("%.17g" % v for v in th) + "\n")
            env = dict(os.environ, OPENBLAS_NUM_THREADS="1", GPSS_REF_KERNEL=a + "+" + b)
            out = subprocess.run([DRIVER, os.path.join(d, "train.txt"), os.path.join(d, "test.txt"), os.path.join(d, "thetas.txt"), "3", d],
                                 capture_output=True, text=True, check=True, env=env)
            r = parse_dump(out.stdout)
            # the reference's own command line on the same file
            cli = os.path.join(ROOT, "oracle", "_ref", "gp_ss_ak")
            tr = subprocess.run([cli, "-v", "3", "-pm", "1", "train", "-k", a, "-k", b, "-kn", "1", "-o", "LBFGS", "-#", "2",
                                 os.path.join(d, "train.txt"), os.path.join(d, "cli_model")], capture_output=True, text=True,
                                stdin=subprocess.DEVNULL, env=dict(env, MALLOC_PERTURB_="255"), cwd=d)
            tag = a + "_" + b
            rec[tag + "_cli_train_stdout"] = np.array(tr.stdout)
            rec[tag + "_cli_rc"] = tr.returncode
            rec[tag + "_cli_model_text"] = np.array(open(os.path.join(d, "cli_model")).read()) if os.path.exists(os.path.join(d, "cli_model")) else np.array("")
            rec["train_file_text"] = np.array(open(os.path.join(d, "train.txt")).read())
        for k in range(len(ths)):
            for key in ("theta", "nlml", "g", "alpha", "K_diag", "K_col17", "mu", "var"):
                rec["%s_%s_%d" % (tag, key, k)] = r["%s_%d" % (key, k)]
        for key in ("probe_kind", "probe_theta", "probe_f", "probe_g", "theta_fit", "nlml_fit"):
            if key in r:
                rec[tag + "_" + key] = r[key]
        rec["Xs"] = r["Xs"]; rec["ys"] = r["ys"]; rec["params"] = r["params"]; rec["Xt"] = r["Xt"]
        print(tag, "nlml", [r["nlml_%d" % k] for k in range(len(ths))], "g0", np.round(r["g_0"].ravel(), 4), "cli rc", tr.returncode, "probes", len(r.get("probe_f", [])))
    np.savez_compressed(os.path.join(HERE, "ref_sum2_n300.npz"), **rec)
