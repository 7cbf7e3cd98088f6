"""The widened configurations at n = 1 000, recorded from the UNMODIFIED reference like
make_ref_golden.py does at n = 300: 4-column rock-type ExpAns, Hyb{Exp, Bias}, Hyb{RBF, Bias}; two evaluations and a
4-iteration LBFGS probe trace each.  Build container only (needs oracle/_ref).

    python tests/golden/make_ref_n1000_widened.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_ref_golden as M          # noqa: E402

if __name__ == "__main__":
    if not os.path.exists(M.DRIVER):
        raise SystemExit("oracle/_ref/ref_driver is not built (needs /root/reference): make -C oracle")
    rng = np.random.default_rng(1000)
    th1 = np.clip(M.THETA0 * rng.uniform(0.8, 1.25, 10), 1e-4, 6.0)
    M.make("ref_rock_n1000.npz", 1000, 11, [M.THETA0, th1], lbfgs_iters=4, rock=True)
    for kernel, th0 in (("Exp", M.THETA0_EXP), ("RBF", M.THETA0_RBF)):
        ths = [th0, np.clip(th0 * rng.uniform(0.8, 1.25, th0.shape[0]), 1e-4, 6.0)]
        M.make("ref_%s_n1000.npz" % kernel.lower(), 1000, 12, ths, lbfgs_iters=4, kernel=kernel)
