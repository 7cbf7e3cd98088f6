"""Distributed-evaluation check, launched with torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py [n ...]
Every rank holds the same data; rank 0 also evaluates on a private single-GPU handle and the results are compared."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gp_ss_ak_b200 as G
from gp_ss_ak_b200 import datagen

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sizes = [int(a) for a in sys.argv[1:]] or [3000, 20000]
base = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])
ok = True
for n in sizes:
    X, y = datagen.drillholes(n, 0)
    Xs, ys, params = datagen.standardise_symmetric(X, y)
    m = G.GpssModel(Xs, ys, device=local)
    ids = [G.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m.dist_init(rank, world, ids[0])
    ref = None
    if rank == 0 and n <= 20000:
        # same pipe for the long-k products as the multi-GPU handle (the int8 pipe unless GPSS_OZAKI_DIST=0 asked for DMMA)
        prev = os.environ.get("GPSS_OZAKI")
        os.environ["GPSS_OZAKI"] = str(m.ozaki_slices())
        prev_bits = os.environ.get("GPSS_OZAKI_BITS")
        if m.ozaki_slices():
            os.environ["GPSS_OZAKI_BITS"] = str(m.ozaki_digit_bits())      # same digits as the multi-GPU handle (7 slices of 8 bits by default)
        ms = G.GpssModel(Xs, ys, device=local)
        if prev is None:
            os.environ.pop("GPSS_OZAKI", None)
        else:
            os.environ["GPSS_OZAKI"] = prev
        if prev_bits is None:
            os.environ.pop("GPSS_OZAKI_BITS", None)
        else:
            os.environ["GPSS_OZAKI_BITS"] = prev_bits
        ms.set_theta(base)
        ref = ms.nlml_grad() + (ms.alpha(),)
        ms.close()
    for rep in range(3):
        th = base * (1 + 0.01 * rep)
        m.set_theta(th)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        L, g = m.nlml_grad()
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        if rep == 0:
            L0, g0, a0 = L, g.copy(), m.alpha()
        if rank == 0:
            print("n %d world %d rep %d nlml %.12f wall %.1f ms device %.1f ms" % (n, world, rep, L, dt * 1e3, m.last_call_ms()), flush=True)
    # phase breakdown of one more evaluation, per rank
    if os.environ.get("GPSS_DIST_PHASES"):
        m.set_profiling(True)
        m.set_theta(base * 1.005)
        m.nlml_grad()
        ph = m.phase_ms()
        m.set_profiling(False)
        allph = [None] * world
        dist.all_gather_object(allph, [float(v) for v in ph[:9]])
        if rank == 0:
            names = ["kbuild", "potrf", "solve", "trtri", "lauum", "grad", "cross", "vargemm", "gatherU"]
            for r_, p_ in enumerate(allph):
                print("   rank %d phases ms: " % r_ + " ".join("%s %.1f" % (nm, v) for nm, v in zip(names, p_) if v > 0), flush=True)
    # all ranks must hold identical results
    t = torch.tensor(np.concatenate([[L0], g0]), device="cuda")
    tl = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(tl, t)
    same = all(torch.equal(tl[0], v) for v in tl)
    if rank == 0:
        print("   identical across ranks:", same, flush=True)
        ok &= same
        if ref is not None:
            eL = abs(L0 - ref[0]) / abs(ref[0]); eg = np.abs(g0 - ref[1]).max() / np.abs(ref[1]).max()
            ea = np.linalg.norm(a0 - ref[2]) / np.linalg.norm(ref[2])
            print("   vs single GPU: nlml rel %.2e  g rel %.2e  alpha rel %.2e" % (eL, eg, ea), flush=True)
            ok &= eL < 1e-11 and eg < 1e-9 and ea < 1e-10
    # sharded prediction: every rank predicts its slice of the test points, slices gathered, one post-processing
    Xt_raw, _ = datagen.drillholes(2000, 5)
    Xt = (Xt_raw - params[1:, 0]) / params[1:, 1]
    m.set_theta(base)
    sums = np.array([np.add.reduce(Xt[:, j].tolist()) for j in range(3)])
    from oracle import gpss_oracle as O
    sums = O.seq_colsum(Xt)
    lo, hi = rank * Xt.shape[0] // world, (rank + 1) * Xt.shape[0] // world
    mu_s, var_s = m.predict_shard(Xt.shape[0], sums, Xt[lo:hi])
    parts = [None] * world
    dist.all_gather_object(parts, (mu_s, var_s))
    if rank == 0:
        mu = np.concatenate([p[0] for p in parts]); var = G.var_postprocess(np.concatenate([p[1] for p in parts]), base[9])
        if n <= 20000:
            ms = G.GpssModel(Xs, ys, device=local); ms.set_theta(base)
            mu1, var1 = ms.predict(Xt); ms.close()
            print("   sharded prediction vs single GPU: mu %.2e var %.2e" % (np.abs(mu - mu1).max(), np.abs(var - var1).max()), flush=True)
            ok &= np.abs(mu - mu1).max() < 1e-9 and np.abs(var - var1).max() < 1e-9
    m.close()
if rank == 0:
    print("DIST CHECK", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
