"""Multi-GPU A/B of run-time switches on ONE replicated handle (torchrun, one rank per GPU): for every setting the environment is changed
on all ranks, two evaluations run and the second is timed.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_sweep.py n "A=1 B=2" "A=3" ...
An empty string is the default configuration."""
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
n = int(sys.argv[1])
settings = sys.argv[2:] or [""]
base = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])
X, y = datagen.drillholes(n, 0)
Xs, ys, _ = datagen.standardise_symmetric(X, y)
m = G.GpssModel(Xs, ys, device=local)
if world > 1:
    ids = [G.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    m.dist_init(rank, world, ids[0])
m.set_theta(base)
L_ref, g_ref = m.nlml_grad()
touched = set()
for st in settings:
    for k in touched:
        os.environ.pop(k, None)
    for kv in st.split():
        k, v = kv.split("=", 1)
        os.environ[k] = v
        touched.add(k)
    ms = []
    for rep in range(3):
        m.set_theta(base * (1 + 0.01 * rep))
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        L, g = m.nlml_grad()
        torch.cuda.synchronize(); dist.barrier()
        ms.append(m.last_call_ms())
        if rep == 0:
            dL, dg = abs(L - L_ref) / abs(L_ref), np.abs(g - g_ref).max() / np.abs(g_ref).max()
    if rank == 0:
        print("n %d world %d [%s]: %.1f / %.1f / %.1f ms   (nlml %.1e, g %.1e from the first evaluation)" % (n, world, st or "default", ms[0], ms[1], ms[2], dL, dg), flush=True)
m.close()
dist.destroy_process_group()
