"""Synthetic 3-D drillhole-style ore-grade data; the reference ships no data.

Harness utility (tests / bench); numpy only.  The generated doubles are written with 17 significant
digits so the reference reader's atof (Control.cpp:65,72) reproduces them exactly.
"""
from __future__ import annotations

import math
import numpy as np


def _rotation(ax_deg, ay_deg, az_deg):
    a, b, c = (math.radians(v) for v in (ax_deg, ay_deg, az_deg))
    Rx = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]])
    Ry = np.array([[math.cos(b), 0, math.sin(b)], [0, 1, 0], [-math.sin(b), 0, math.cos(b)]])
    Rz = np.array([[math.cos(c), -math.sin(c), 0], [math.sin(c), math.cos(c), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _field(P, rng_field):
    """f(x) = sum_k a_k exp(-|Lambda^(1/2) R (x - c_k)|), 32 anisotropic bumps."""
    nb = 32
    a = rng_field.standard_normal(nb)
    c = rng_field.uniform(0.0, 1.0, size=(nb, 3)) * np.array([1000.0, 1000.0, 400.0])
    R = _rotation(30.0, 20.0, 10.0)
    inv_ranges = 1.0 / np.array([300.0, 150.0, 50.0])
    f = np.zeros(P.shape[0])
    for k in range(nb):
        d = (P - c[k]) @ R.T
        d *= inv_ranges
        f += a[k] * np.exp(-np.sqrt((d * d).sum(axis=1)))
    return f


def drillholes(n, seed=0):
    """n samples along ceil(n/100) jittered-grid drillholes in a 1000 x 1000 x 400 m domain.
    Returns X (n,3) raw metres and y (n,) grades (log-normal-ish, positive)."""
    rng = np.random.default_rng(seed)
    H = int(math.ceil(n / 100))
    side = int(math.ceil(math.sqrt(H)))
    pitch = 1000.0 / side
    pts = []
    for h in range(H):
        gx, gy = h % side, h // side
        cx = (gx + 0.5) * pitch + rng.uniform(-0.4, 0.4) * pitch
        cy = (gy + 0.5) * pitch + rng.uniform(-0.4, 0.4) * pitch
        az = rng.uniform(0.0, 2 * math.pi)
        dip = math.radians(rng.uniform(60.0, 90.0))
        direction = np.array([math.cos(dip) * math.cos(az), math.cos(dip) * math.sin(az), -math.sin(dip)])
        depth = np.arange(100) * 2.0
        p = np.array([cx, cy, 400.0]) + depth[:, None] * direction[None, :]
        p += rng.normal(0.0, 0.05, size=p.shape)
        pts.append(p)
    X = np.concatenate(pts, axis=0)[:n]
    f = _field(X, np.random.default_rng(seed + 1000003))
    y = np.exp(0.5 * f + 0.1 * rng.standard_normal(n))
    return np.ascontiguousarray(X), np.ascontiguousarray(y)


def with_rock_column(P, seed=0):
    """Appends a 4th input column: a rock-type code 1..4 (lithology bands that dip across the domain, a deterministic function
    of position), the input of the reference's 4-column ExpAns branch (InversewidthR, Kernel.cpp:872-878)."""
    t = 0.004 * P[:, 0] + 0.003 * P[:, 1] + 0.006 * P[:, 2] + 0.37 * (seed % 7)
    code = 1.0 + np.floor(np.mod(t, 4.0))
    return np.ascontiguousarray(np.concatenate([P, code[:, None]], axis=1))


def block_model(nx, ny, nz, lo, hi):
    """Block-model centroids on a regular nx x ny x nz grid over the bounding box [lo, hi]."""
    ax = [lo[d] + (np.arange(k) + 0.5) * (hi[d] - lo[d]) / k for d, k in enumerate((nx, ny, nz))]
    G = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3)
    return np.ascontiguousarray(G)


def grade_at(P, seed=0):
    """Noise-free grade of the seed's field at arbitrary points (for test files' last column)."""
    return np.exp(0.5 * _field(P, np.random.default_rng(seed + 1000003)))


def standardise_symmetric(X, y):
    """Numerically the same result as Control::prep_symmetric in train mode (Control.cpp:299-324): the three
    spatial columns share one centre/half-range from the global min/max.  Returns Xs, ys, params((D+1) x 2)."""
    D = X.shape[1]
    params = np.zeros((D + 1, 2))
    params[0] = (0.5 * (y.max() + y.min()), 0.5 * (y.max() - y.min()))
    xmax, xmin = X.max(), X.min()
    for j in range(min(3, D)):
        params[j + 1] = (0.5 * (xmax + xmin), 0.5 * (xmax - xmin))
    for j in range(3, D):
        params[j + 1] = (0.5 * (X[:, j].max() + X[:, j].min()), 0.5 * (X[:, j].max() - X[:, j].min()))
    Xs = (X - params[1:, 0]) / params[1:, 1]
    ys = (y - params[0, 0]) / params[0, 1]
    return np.ascontiguousarray(Xs), np.ascontiguousarray(ys), params


def write_data_file(path, X, y):
    """Tab-delimited 'x y z grade' lines, 17 significant digits."""
    with open(path, "w") as f:
        for i in range(X.shape[0]):
            f.write("\t".join("%.17g" % v for v in X[i]) + "\t%.17g\n" % y[i])
