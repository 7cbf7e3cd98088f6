"""BASELINE.json configs[0] -- "ExpAns train, LBFGS, synthetic 3D ore-grade set n=2,000 (./gp_ss_ak -v 3 -pm 1 train -k ExpAns
-kn 1 -o LBFGS), runs on CPU" -- recorded from the UNMODIFIED reference (oracle/_ref, see make_ref_golden.py for the fixture
layout): two single evaluations, a bounded LBFGS probe trace and the reference command line's own train + test outputs.
Build container only (needs oracle/_ref); several minutes of CPU (one BLAS thread, -O0 reference).

    python tests/golden/make_ref_n2000.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_ref_golden as M          # noqa: E402

if __name__ == "__main__":
    if not os.path.exists(M.DRIVER):
        raise SystemExit("oracle/_ref/ref_driver is not built (needs /root/reference): make -C oracle")
    rng = np.random.default_rng(2000)
    th1 = np.clip(M.THETA0 * rng.uniform(0.8, 1.25, 10), 1e-4, 6.0)
    M.make("ref_n2000.npz", 2000, 0, [M.THETA0, th1], lbfgs_iters=4, n_test=200, n_coincident=20, cli_iters=2)
