"""One rank of the world_size-2 CPU (`gloo`) job run by tests/test_dist_cpu.py.

(1) Block-column Cholesky: executes the SAME per-rank operation list the CUDA driver executes (gpss_dist_potrf_schedule,
    host logic of potrf_blocked in gpss_capi.cu) with numpy standing in for the DMMA kernels and a gloo broadcast for the
    NCCL panel broadcast; the factor every rank ends up with must equal numpy's Cholesky.
(2) Sharded prediction: test points split over the ranks with the GLOBAL Mahalanobis centre, raw variances gathered, the
    reference's whole-vector post-processing applied once (gpss_var_postprocess) -- must equal the unsharded oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen
from oracle import gpss_oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)

# ---- (1) the block-column algorithm on a small SPD matrix; block width 8 stands in for the 512 of the CUDA path ----
W, nblk = 8, 11
n = W * nblk - 3                                   # ragged last block column
rng = np.random.default_rng(7)
M = rng.standard_normal((n, n))
Afull = M @ M.T + n * np.eye(n)
A = np.tril(Afull).copy()                          # every rank starts from the same matrix ("K build")
for j in range(nblk):                              # ... but only OWN block columns are trusted: poison the others
    if j % world != rank:
        A[:, j * W:(j + 1) * W] = np.nan
have = set()
for kind, col, pbeg, pcnt, root, stream in G.dist_potrf_schedule(nblk, world, rank):
    c0, c1 = col * W, min((col + 1) * W, n)
    if kind in (1, 4):                             # trailing update of MY block column with panels pbeg .. pbeg+pcnt-1
        assert all(p in have for p in range(pbeg, pbeg + pcnt))
        k0, k1 = pbeg * W, (pbeg + pcnt) * W
        A[c0:, c0:c1] -= A[c0:, k0:k1] @ A[c0:c1, k0:k1].T
    elif kind == 2:                                # factor the block column
        Ld = np.linalg.cholesky(np.tril(A[c0:c1, c0:c1]) + np.tril(A[c0:c1, c0:c1], -1).T)
        A[c0:c1, c0:c1] = Ld
        A[c1:, c0:c1] = np.linalg.solve(Ld, A[c1:, c0:c1].T).T
    elif kind == 3:                                # panel broadcast
        t = torch.from_numpy(np.ascontiguousarray(A[c0:, c0:c1]))
        dist.broadcast(t, src=root)
        A[c0:, c0:c1] = t.numpy()
        have.add(col)
L = np.tril(A)
Lref = np.linalg.cholesky(Afull)
assert np.isfinite(L).all()
assert np.abs(L - Lref).max() < 1e-10 * np.abs(Lref).max(), np.abs(L - Lref).max()

# ---- (2) sharded prediction flow ----
X, y = datagen.drillholes(160, 2)
Xs, ys, params = datagen.standardise_symmetric(X, y)
Xt_raw, _ = datagen.drillholes(50, 9)
Xt = (Xt_raw - params[1:, 0]) / params[1:, 1]
Xt[:5] = Xs[:5]                                    # coincident points
gp = O.OracleGP(Xs, ys, O.THETA0, literal=False)
mu_whole, var_whole = gp.predict(Xt)
sums = O.seq_colsum(Xt)
c = O.centre(Xs, Xt, sums2=sums)                   # the centre over the training set and ALL test points (Kernel.cpp:1391)
lo, hi = rank * Xt.shape[0] // world, (rank + 1) * Xt.shape[0] // world
kX, _ = O.compute_K(Xs, Xt[lo:hi], O.THETA0, c=c)
gp.update_alpha()
gp.log_likelihood()
Wh = np.sqrt(gp.d2lp)
V = gp._solve_chol(gp.Lchol, kX * Wh[:, None]) * Wh[:, None] * kX
parts = [None] * world
dist.all_gather_object(parts, (kX.T @ gp.Alpha, (O.THETA0[6] ** 2 + O.THETA0[8]) - V.sum(axis=0)))
mu = np.concatenate([p[0] for p in parts])
var = G.var_postprocess(np.concatenate([p[1] for p in parts]), O.THETA0[9])      # host function of libgpss.so, once, rank-independent
assert np.abs(mu - mu_whole).max() < 1e-12 and np.abs(var - var_whole).max() < 1e-12
assert var[0] == O.THETA0[9]                       # the reference's index quirk acts on the GATHERED vector, not per shard

dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("WORKER OK")
