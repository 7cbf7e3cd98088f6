"""Partitioned-storage check (BASELINE config 5), launched with torchrun, one rank per GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/part_check.py [n ...]
Every rank holds the same data; the factor lives as block columns spread over the ranks.  For n <= 20000 rank 0 also evaluates
on a private single-GPU handle and the results are compared; at any n the solve is checked against the definition:
   || K alpha + sn2 alpha - y || / || y ||      with K alpha recomputed from coordinates (kmatvec), independent of the factor."""
import json
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
    ids = [G.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m = G.GpssModel(Xs, ys, device=local, partitioned=(rank, world, ids[0]))
    n_pad = m.padded_n()
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
        ref = (ms.nlml(), ms.alpha(), ms.yhat(), ms.predict(Xs[:512] * 0.99, want_var=False)[0], ms.nlml_grad()[1],
               ms.predict(Xs[:700] * 0.99, want_var=True))
        ms.close()
    times = []
    for rep in range(3 if n <= 60000 else 1):
        th = base * (1 + 0.01 * rep)
        m.set_theta(th)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        L = m.nlml()
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        times.append(m.last_call_ms())
        if rank == 0:
            print("n %d world %d rep %d nlml %.12f wall %.1f ms device %.1f ms" % (n, world, rep, L, dt * 1e3, m.last_call_ms()), flush=True)
    m.set_profiling(True)
    m.set_theta(base)
    L0 = m.nlml()
    ph = m.phase_ms()
    m.set_profiling(False)
    a0, f0 = m.alpha(), m.yhat()
    mu0, _ = m.predict(Xs[:512] * 0.99, want_var=False)
    resid = np.linalg.norm(f0 + base[9] * a0 - ys) / np.linalg.norm(ys)
    allph = [None] * world
    dist.all_gather_object(allph, [float(v) for v in ph[:3]])
    t = torch.tensor(np.concatenate([[L0], a0[:4096], mu0]), device="cuda")
    tl = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(tl, t)
    same = all(torch.equal(tl[0], v) for v in tl)
    if rank == 0:
        chol_tf = (float(n_pad) ** 3 / 3) / (max(p[1] for p in allph) * 1e-3) * 1e-12
        print("   phases ms per rank (kbuild, potrf, solve): " + " | ".join("%.1f %.1f %.1f" % tuple(p) for p in allph), flush=True)
        print("   identical across ranks: %s   residual |K a + sn2 a - y|/|y| = %.3e   Cholesky %.1f TFLOP/s aggregate (%.1f per GPU)"
              % (same, resid, chol_tf, chol_tf / world), flush=True)
        ok &= same and resid < 1e-8
        if ref is not None:
            eL = abs(L0 - ref[0]) / abs(ref[0]); ea = np.linalg.norm(a0 - ref[1]) / np.linalg.norm(ref[1])
            ef = np.abs(f0 - ref[2]).max(); em = np.abs(mu0 - ref[3]).max()
            print("   vs single GPU: nlml rel %.2e  alpha rel %.2e  yhat abs %.2e  mean abs %.2e" % (eL, ea, ef, em), flush=True)
            ok &= eL < 1e-11 and ea < 1e-9 and ef < 1e-9 and em < 1e-9
        print("PART_RESULT " + json.dumps({"n": n, "n_pad": n_pad, "world": world, "nlml": L0, "nlml_ms": float(np.min(times)),
                                           "potrf_ms": max(p[1] for p in allph), "solve_ms": max(p[2] for p in allph),
                                           "cholesky_tflops": chol_tf, "residual": resid, "identical": bool(same)}), flush=True)
    # the gradient: U = L^-T as cyclic block rows, B^-1 consumed strip by strip (gpss_nlml_grad on a partitioned handle)
    if os.environ.get("GPSS_PART_GRAD", "1") != "0":
        m.set_profiling(True)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        Lg, g = m.nlml_grad()                      # the objective at `base` is cached: this times the inverse + gradient only
        torch.cuda.synchronize(); dist.barrier()
        dtg = time.perf_counter() - t0
        phg = m.phase_ms()
        m.set_profiling(False)
        tg = torch.tensor(g, device="cuda")
        tgl = [torch.zeros_like(tg) for _ in range(world)]
        dist.all_gather(tgl, tg)
        same_g = all(torch.equal(tgl[0], v) for v in tgl)
        if rank == 0:
            print("   gradient: %.1f ms (inverse %.1f, B^-1 strips + reductions %.1f)  identical across ranks: %s" % (dtg * 1e3, phg[3], phg[4], same_g), flush=True)
            print("   g = " + " ".join("%.6e" % v for v in g), flush=True)
            ok &= same_g and bool(np.all(np.isfinite(g))) and Lg == L0
            if ref is not None:
                eg = np.abs(g - ref[4]).max() / np.abs(ref[4]).max()
                print("   vs single GPU: g rel %.2e   (ref g = %s)" % (eg, " ".join("%.6e" % v for v in ref[4])), flush=True)
                ok &= eg < 1e-9
            print("PART_GRAD_RESULT " + json.dumps({"n": n, "world": world, "grad_ms": dtg * 1e3, "trtri_ms": float(phg[3]), "binv_grad_ms": float(phg[4])}), flush=True)
    # predictive variance from the partitioned factor (variance_partitioned: strips of U = L^-T broadcast, k = 512 products per strip);
    # a collective over the same test points on every rank
    if os.environ.get("GPSS_PART_VAR", "1") != "0":
        mt = 700 if n <= 60000 else 4096
        Xv = Xs[:mt] * 0.99
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        mu_v, var_v = m.predict(Xv, want_var=True)
        torch.cuda.synchronize(); dist.barrier()
        dtv = time.perf_counter() - t0
        tv = torch.tensor(np.concatenate([mu_v, var_v]), device="cuda")
        tvl = [torch.zeros_like(tv) for _ in range(world)]
        dist.all_gather(tvl, tv)
        same_v = all(torch.equal(tvl[0], v) for v in tvl)
        if rank == 0:
            print("   predictive mean + variance of %d points: %.1f ms (device %.1f ms)  identical across ranks: %s  var in [%.4g, %.4g]"
                  % (mt, dtv * 1e3, m.last_call_ms(), same_v, var_v.min(), var_v.max()), flush=True)
            ok &= same_v and bool(np.all(np.isfinite(var_v)))
            if ref is not None:
                emu = np.abs(mu_v - ref[5][0]).max(); ev = np.abs(var_v - ref[5][1]).max()
                print("   vs single GPU: mean abs %.2e  variance abs %.2e" % (emu, ev), flush=True)
                ok &= emu < 1e-9 and ev < 1e-9
            print("PART_VAR_RESULT " + json.dumps({"n": n, "world": world, "points": mt, "ms": dtv * 1e3}), flush=True)
    m.close()
if rank == 0:
    print("PART CHECK", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
