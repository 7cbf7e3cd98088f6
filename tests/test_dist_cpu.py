"""CPU tests of the multi-GPU host logic (no device): the distributed Cholesky's per-rank operation lists replayed for all
ranks, the balanced row partitions, and -- as a real world_size-2 `gloo` job -- the shard / gather / post-process flow
of sharded prediction and the numerics of the block-column algorithm itself (numpy stands in for the DMMA kernels)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WAIT_SIDE, UPDATE_MAIN, FACTOR, BCAST, UPDATE_SIDE = range(5)


def _replay(gpss, nblk, world):
    """Execute the op lists of all ranks in lock step on the broadcasts; returns per-rank, per-column applied panels."""
    ops = [gpss.dist_potrf_schedule(nblk, world, r) for r in range(world)]
    pos = [0] * world
    have = [set() for _ in range(world)]                  # panels complete on the rank
    applied = [dict() for _ in range(world)]              # column -> list of panels applied (own columns only)
    side_pending = [dict() for _ in range(world)]         # column -> panels applied on a side stream, not yet waited for
    factored = [set() for _ in range(world)]
    bcasts = 0
    while any(pos[r] < len(ops[r]) for r in range(world)):
        progressed = False
        # run every rank up to its next broadcast
        for r in range(world):
            while pos[r] < len(ops[r]) and ops[r][pos[r]][0] != BCAST:
                kind, col, pbeg, pcnt, root, stream = ops[r][pos[r]]
                assert col % world == r, "rank %d touches column %d it does not own" % (r, col)
                if kind == WAIT_SIDE:
                    side_pending[r].pop(col, None)
                elif kind in (UPDATE_MAIN, UPDATE_SIDE):
                    panels = list(range(pbeg, pbeg + pcnt))
                    assert all(p in have[r] for p in panels), "rank %d applies a panel it does not have yet: %r" % (r, panels)
                    assert all(p < col for p in panels)
                    applied[r].setdefault(col, []).extend(panels)
                    if kind == UPDATE_SIDE:
                        side_pending[r].setdefault(col, []).extend(panels)
                        assert stream in (0, 1)
                elif kind == FACTOR:
                    assert col not in side_pending[r], "column %d factored before its side-stream updates were waited for" % col
                    assert sorted(applied[r].get(col, [])) == list(range(col)), (r, col, applied[r].get(col))
                    factored[r].add(col)
                pos[r] += 1
                progressed = True
        # all ranks must now be at the SAME broadcast (NCCL collectives are matched by issue order)
        heads = [ops[r][pos[r]] if pos[r] < len(ops[r]) else None for r in range(world)]
        if all(h is None for h in heads):
            break
        assert all(h is not None and h[0] == BCAST for h in heads)
        assert len({(h[1], h[4]) for h in heads}) == 1, "ranks disagree on the broadcast: %r" % (heads,)
        col, root = heads[0][1], heads[0][4]
        assert root == col % world and col in factored[root], "column %d broadcast before its owner factored it" % col
        for r in range(world):
            have[r].add(col)
            pos[r] += 1
        bcasts += 1
        progressed = True
        assert progressed
    assert bcasts == nblk
    for r in range(world):
        assert have[r] == set(range(nblk))
    return ops, applied


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("nblk", [1, 2, 5, 9, 40, 98])
def test_distributed_cholesky_schedule_is_complete_and_ordered(gpss, nblk, world):
    ops, applied = _replay(gpss, nblk, world)
    for r in range(world):
        for col, panels in applied[r].items():
            assert len(panels) == len(set(panels)) == col            # every earlier panel exactly once


def test_schedule_keeps_bulk_work_off_the_critical_path(gpss):
    """At world 8 the main stream of a rank applies exactly ONE panel (k = 512) per own column; everything else is queued on
    the side streams at least one broadcast earlier, and the long-k chunk as early as the rank's previous own column."""
    nblk, world = 98, 8
    for r in range(world):
        ops = gpss.dist_potrf_schedule(nblk, world, r)
        main_panels = sum(o[3] for o in ops if o[0] == UPDATE_MAIN)
        own = [j for j in range(nblk) if j % world == r]
        assert main_panels == len([j for j in own if j >= 1])
        for i, o in enumerate(ops):
            if o[0] == UPDATE_SIDE and o[3] > 1:                     # chunk A(j): issued right after broadcast j - world
                assert ops[i - 1][0] == BCAST and ops[i - 1][1] == o[1] - world and o[2] == 0 and o[3] == o[1] - world + 1


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_balanced_row_partitions(gpss, world):
    n_pad = 50048
    for kind in (0, 1):
        b = gpss.dist_partition(n_pad, world, kind)
        assert b[0] == 0 and b[-1] == n_pad and all(x % 128 == 0 for x in b) and all(b[i] <= b[i + 1] for i in range(world))
        x = np.array(b, dtype=float)
        n = float(n_pad)
        # work integrals of the two partitions (gpss_capi.cu balanced_rows): equal shares within the 128-row rounding
        w = (n - x[:-1]) ** 3 - (n - x[1:]) ** 3 if kind == 0 else (n * x[1:] ** 2 / 2 - x[1:] ** 3 / 3) - (n * x[:-1] ** 2 / 2 - x[:-1] ** 3 / 3)
        assert w.max() / w.mean() < 1.05


def _inverse_step_model(a_t, t, ntile, tile_cols=8, slots=148, tpb=4):
    """The cost model behind kind 2 (gpss_potrf.cuh: inv_slice_cost), restated: per block column the CTAs of the rows [a_t, a_t + t) above
    it (tile_cols per row, longest k-range first) run in waves of `slots`, each wave as long as its first CTA."""
    cost = 0.0
    for J0t in range(tpb, ntile, tpb):
        r1 = min(a_t + t, J0t)
        if r1 <= a_t:
            continue
        jobs = (r1 - a_t) * tile_cols
        w = 0
        while w * slots < jobs:
            cost += J0t - (a_t + (w * slots) // tile_cols)
            w += 1
    return cost


@pytest.mark.parametrize("world", [2, 4, 8])
def test_wave_aware_row_partition_of_the_int8_inverse(gpss, world):
    """kind 2 (rows of U = L^-T on the int8 pipe, one CTA per SM): a valid partition whose largest rank cost under the model is not above the
    flop-balanced partition's (kind 0), and clearly below it at 8 ranks, where kind 0 hands ranks 1-4 slices of 19-24 tile rows = two waves per step."""
    n_pad = 50048
    ntile = n_pad // 128
    cost = {}
    for kind in (0, 2):
        b = gpss.dist_partition(n_pad, world, kind)
        assert b[0] == 0 and b[-1] == n_pad and all(x % 128 == 0 for x in b) and all(b[i] <= b[i + 1] for i in range(world))
        cost[kind] = max(_inverse_step_model(b[k] // 128, (b[k + 1] - b[k]) // 128, ntile) for k in range(world))
    assert cost[2] <= cost[0] * 1.0001
    if world == 8:
        assert cost[2] < 0.85 * cost[0]


@pytest.mark.parametrize("n_pad,world", [(50048, 8), (50048, 2), (20096, 4), (3072, 2), (1152, 3)])
def test_exchange_layout_of_the_inverse_slices(gpss, n_pad, world):
    """allgather_U sends a rank's rows of U = L^-T without the 512-wide diagonal blocks and the zeros below them (uslice_copy_kernel).  The
    packed layout must be a bijection onto [0, count): offsets increase by exactly the column lengths, the last column ends at the allocated
    size, a column keeps exactly its rows strictly above its own diagonal block, and the slices of all ranks together cover every entry of
    the strict block-upper triangle once."""
    b = gpss.dist_partition(n_pad, world, 0)
    covered = 0
    for k in range(world):
        r0, rows = b[k], b[k + 1] - b[k]
        off, lens, count = gpss.dist_uslice_layout(n_pad, r0, rows)
        off, lens = np.array(off), np.array(lens)
        j = np.arange(r0, n_pad)
        expect = np.clip((j // 512) * 512 - r0, 0, rows)
        assert np.array_equal(lens, expect)
        assert off[0] == 0 and np.array_equal(np.diff(off), lens[:-1]) and off[-1] + lens[-1] == count
        covered += count
    jj = np.arange(n_pad)
    assert covered == int(((jj // 512) * 512).sum())          # rows strictly above the diagonal block of every column


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world2_gloo_block_column_cholesky_and_sharded_prediction(gpss):
    """torchrun-style world_size-2 job on CPU (gloo): see tests/dist_worker.py."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), WORLD_SIZE="2", OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
    assert "WORKER OK" in outs[0]


@pytest.mark.parametrize("P,nb", [(2, 5), (3, 7), (8, 11)])
def test_partitioned_inverse_algorithm_and_index_maps(P, nb):
    """numpy replay of trtri_partitioned / gradient_partitioned (gpss_capi.cu) with the SAME index maps the CUDA path uses:
    block columns of L and block rows of U = L^-T owned cyclically and packed (local block q <-> global q P + r), the L row
    strip assembled from an owner-major all-gather (segment stride jc blocks, block k at segment k % P, slot k / P -- the
    formula of order_rowstrip_kernel), the prefix / suffix row counts, and B^-1 produced strip by strip.  Must reproduce
    inv(L)^T and the lower triangle of inv(B)."""
    W = 4
    n = W * nb
    rng = np.random.default_rng(P * 100 + nb)
    M = rng.standard_normal((n, n))
    B = M @ M.T + n * np.eye(n)
    L = np.linalg.cholesky(B)
    own = lambda j: j % P
    Lloc = [np.concatenate([L[:, j * W:(j + 1) * W] for j in range(nb) if own(j) == r] or [np.zeros((n, 0))], axis=1) for r in range(P)]
    nq = [Lloc[r].shape[1] // W for r in range(P)]
    Uloc = [np.zeros((max(nq[r], 1) * W, n)) for r in range(P)]
    for J in range(nb):
        J0, o = J * W, own(J)
        Wjj = np.linalg.inv(Lloc[o][J0:J0 + W, (J // P) * W:(J // P + 1) * W])
        if J > 0:
            jc = (J + P - 1) // P                                               # blocks per all-gather segment
            gathered = np.full((P * jc, W, W), np.nan)
            for r in range(P):
                cnt = len([q for q in range(nq[r]) if q * P + r < J])
                for b in range(cnt):                                            # pack_rowstrip_kernel: my first cnt block columns, rows J0..
                    gathered[r * jc + b] = Lloc[r][J0:J0 + W, b * W:(b + 1) * W]
            Lrow = np.concatenate([gathered[(k % P) * jc + k // P] for k in range(J)], axis=1)   # order_rowstrip_kernel
            assert np.isfinite(Lrow).all() and np.array_equal(Lrow, L[J0:J0 + W, :J0])
        for r in range(P):
            cnt = len([q for q in range(nq[r]) if q * P + r < J])
            if cnt and J > 0:
                T = Uloc[r][:cnt * W, :J0] @ Lrow.T
                Uloc[r][:cnt * W, J0:J0 + W] = -T @ Wjj.T
            if r == o:
                Uloc[r][(J // P) * W:(J // P + 1) * W, J0:J0 + W] = Wjj.T
    Uref = np.linalg.inv(L).T
    for r in range(P):
        for q in range(nq[r]):
            I = q * P + r
            assert np.allclose(Uloc[r][q * W:(q + 1) * W], Uref[I * W:(I + 1) * W], atol=1e-12)
    Q = np.zeros((n, n))
    for J in range(nb):
        J0, o = J * W, own(J)
        strip = Uloc[o][(J // P) * W:(J // P + 1) * W, J0:]
        for r in range(P):
            q0 = len([q for q in range(nq[r]) if q * P + r < J])
            if q0 < nq[r]:
                Qs = Uloc[r][q0 * W:nq[r] * W, J0:] @ strip.T
                for q in range(q0, nq[r]):
                    Q[(q * P + r) * W:(q * P + r + 1) * W, J0:J0 + W] = Qs[(q - q0) * W:(q - q0 + 1) * W]
    assert np.allclose(np.tril(Q), np.tril(np.linalg.inv(B)), atol=1e-12)
