#!/usr/bin/env python
"""bench.py -- headline benchmark of the exact-GP hot path: ExpAns LML+gradient evaluations/s at n=50,000.

One "step" = set_GP_Pars(theta_k) followed by Grad_Values(g) (GP_Utils.cpp:130,1171) on synthetic 3-D
drillhole-style data, i.e. K build + FP64 Cholesky + alpha + objective + B^-1 (trtri+lauum) + fused gradient
reductions.  Every step uses a different theta, so nothing is cached between steps; K (20 GB), L and B^-1 do
not fit in the 126 MB L2, so no L2 flush is needed between iterations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 50000] [--impl reference]

N > 1 (launched by torchrun, one rank per GPU): ONE evaluation is partitioned over all GPUs (DESIGN.md
section 7: block-column-cyclic Cholesky with an NCCL panel broadcast, row-sliced inverse, all-reduced
gradient partials); value = evaluations / max-over-ranks device time, scaling "strong".  --mode replicas
runs one independent evaluation stream per GPU instead (weak scaling, no collective).
The line also carries the prediction leg (metric iii): mean + variance of --pred-m block-model centroids
per GPU against the same n-point model, test points split over the GPUs.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_HEADLINE = 50000
# DRAM bytes (read + write) of ONE launch of the dominant kernel from the committed ncu --set full capture
# (profiles/r01_gemm_ws_ncu_full_summary.txt): the B^-1 = U U^T launch at n_pad = 20096 (76.7 ms, 96.7 % DMMA-pipe activity);
# its algorithmic operand bytes are 2 x 1.6 GB read + 1.6 GB written.
GEMM_TRAFFIC_BYTES = 24.010501e9 + 1.599763e9
GEMM_TRAFFIC_NOTE = ("dram__bytes_read.sum + dram__bytes_write.sum of the B^-1 = U U^T launch at n_pad = 20096 (one launch, 76.7 ms); ~5x its "
                     "algorithmic 4.8 GB because operand slabs are re-fetched per wave through L2 (hit rate 83 %), yet only 4 % of the HBM peak: "
                     "the kernel runs at 96.7 % DMMA-pipe activity (profiles/r01_gemm_ws_ncu_full_summary.txt)")
FP64_PEAK_FALLBACK_TFLOPS = 37.13   # profiles/r01_fp64_peak_microbench.txt (DMMA.8x8x4 register loop on this pool's B200)
BF16_SUSTAINED_FALLBACK_TFLOPS = 1400.0   # /opt/skills/guides/B200_PROFILING.md: "sustained it settles near 1.4 PFLOP/s" (used "of fallback")


INT8_PEAK_FALLBACK_TOPS = 3796.9   # profiles/r02_int8_peak_microbench.txt (sustained, N = 256 shape) -- used only if the live measurement fails

# DRAM bytes (read + write) of the B^-1 = U U^T launch of oz_gemm_kernel at n_pad = 50 048, from the ncu capture of the bench command
# with the 7 x 8-bit default: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:oz_gemm_kernel` over
# scripts/profile_target.py 50000 (scripts/gpu_job_r2_m.sh; profiles/r02_oz_gemm_dram_per_launch_n50k.txt, r02_oz_gemm_dram_n50k_7x8bit.csv.gz):
# 360.9 GB read + 17.04 GB written in 351 ms
OZ_LAUUM_TRAFFIC_BYTES = 360.9e9 + 17.04e9


def oz_traffic_note(oz_s, n_pad):
    """Where roofline.traffic of the int8 kernel comes from, and the algorithmic bytes beside it."""
    return ("dram__bytes_read.sum + dram__bytes_write.sum of the B^-1 = U U^T launch at n_pad = 50048 from an ncu pass over one evaluation "
            "(profiles/r02_oz_gemm_dram_per_launch_n50k.txt: ~20x the algorithmic bytes, at 1.08 TB/s = 16 %% of the HBM peak -- every A row panel is "
            "re-read once per super-column of 8 tile columns, every B column panel once per wave); algorithmic: the digit planes of U (upper triangle, "
            "%d planes) read once + the lower triangle of B^-1 written = %.1f GB" % (oz_s, (oz_s * n_pad * n_pad / 2 + 4.0 * n_pad * n_pad) / 1e9))


def int8_tensor_peak(device):
    """int8 tensor-pipe micro-peak in TOP/s measured live (gpss_measure_int8_peak: tcgen05.mma kind::i8 M 128 N 256 K 32 from resident
    shared-memory operands on all SMs).  MEASURED_PEAKS.json has no int8 entry.  The kernel is timed inside a long step, so the
    SUSTAINED figure (~1 s back to back, under the power cap) is the denominator; the burst figure is reported beside it."""
    import gp_ss_ak_b200 as G
    try:
        burst, sust = G.measure_int8_peak(device)
        return sust, burst, "gpss_measure_int8_peak, measured live in this run (sustained over ~1 s; burst = best 6 ms launch); MEASURED_PEAKS.json has no int8 entry"
    except Exception as e:   # noqa: BLE001
        return INT8_PEAK_FALLBACK_TOPS, None, "profiles/r02_int8_peak_microbench.txt (live measurement failed: %s)" % e


def int8_macs_lauum(n_pad):
    """int8 MACs per slice pair of the ONE oz_gemm_kernel launch that forms B^-1 = U U^T (lower tiles, k from the row tile's first row)."""
    mac = 0
    for tm in range(n_pad // 128):
        mac += 128 * 64 * min(n_pad // 64, (tm * 128 + 127) // 64 + 1) * (n_pad - tm * 128)
    return mac


def int8_macs(n_pad, nbo=512):
    """int8 multiply-accumulates PER SLICE PAIR that oz_gemm_kernel executes in one LML+gradient evaluation on one GPU, counted tile by
    tile as the kernel runs them (128 x 64 tiles, whole tiles on the diagonal, triangular k-ranges): the look-ahead updates U1 of the
    Cholesky, the bulk product of the triangular inverse, B^-1 = U U^T (gpss_potrf.cuh / gpss_inverse.cuh).  98.5 % of n_pad^3 / 2 at
    n = 50 000; the rest (k = 512 updates, panels, right factors) stays on the FP64 DMMA pipe."""
    mac = 0
    T = (n_pad + nbo - 1) // nbo
    for t in range(1, T - 1):                                  # U1(t + 1): rows >= T1, block column t + 1, k = 0 .. T0
        T0, T1 = t * nbo, (t + 1) * nbo
        nb1 = min(nbo, n_pad - T1)
        for tm in range((n_pad - T1) // 128):
            mac += 128 * 64 * min(nb1 // 64, (tm * 128 + 127) // 64 + 1) * T0
    for t in range(1, T):                                      # T = U[0:J0, 0:J0] L[J, 0:J0]^T, k from each row tile's first row
        J0 = t * nbo
        nbj = min(nbo, n_pad - J0)
        for tm in range(J0 // 128):
            mac += 128 * nbj * (J0 - tm * 128)
    for tm in range(n_pad // 128):                             # B^-1 = U U^T, lower tiles, k from the row tile's first row
        mac += 128 * 64 * min(n_pad // 64, (tm * 128 + 127) // 64 + 1) * (n_pad - tm * 128)
    return mac


def theta_probe(k):
    """Deterministic theta probes around the reference's initial vector (Kernel.cpp:763-773, GP_Utils.cpp:43)."""
    base = np.array([np.pi / 3.1, 1.5, np.pi / 3.1, 1.5, np.pi / 3.1, 1.3, 0.9, 0.6, 0.2, 0.016])
    rng = np.random.default_rng(1000 + k)
    th = base * rng.uniform(0.9, 1.1, size=10)
    return np.clip(th, 1e-4, 6.0)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def cpu_reference_times(n_sample, warmup, steps):
    """Seconds per LML+gradient evaluation (set_GP_Pars + Grad_Values, a new theta every step) of the reference's own
    CPU implementation on the host cores.  kind "reference": the UNMODIFIED reference classes, built by oracle/Makefile
    into oracle/_ref/ref_driver with the reference's own flags (no optimisation, make_linux:19) and OpenBLAS on all
    cores; kind "port": the numpy/scipy restatement (oracle/gpss_oracle.py) when that binary is not present."""
    from gp_ss_ak_b200 import datagen
    X, y = datagen.drillholes(n_sample, 0)
    if os.path.exists(REF_DRIVER):
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            datagen.write_data_file(os.path.join(d, "train.txt"), X, y)
            out = subprocess.run([REF_DRIVER, "--time", os.path.join(d, "train.txt"), d, str(warmup), str(steps)],
                                 capture_output=True, text=True, check=True)
        secs = [float(l.split()[4]) for l in out.stdout.splitlines() if l.startswith("eval") and " timed " in l]
        return secs, "reference"
    from oracle import gpss_oracle as O
    Xs, ys, _ = datagen.standardise_symmetric(X, y)
    secs = []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        O.nlml_and_grad(Xs, ys, O.THETA0 * (1.0 + 0.01 * ((k % 7) - 3)), dist="blas", literal=True)
        if k >= warmup:
            secs.append(time.perf_counter() - t0)
    return secs, "port"


REF_SIZES = (750, 1500, 3000)          # reference-literal samples (the compiled reference; ~25 s of CPU work in all)
LEAN_SIZES = (5000, 10000)             # lean samples (1 dpotrf + 1 dpotri on all cores; ~15 s); --cpu-lean-max 20000 adds a third


def cpu_baseline_block(n, ref_samples, lean_sizes, cores):
    """the CPU path on the box's host cores in two forms, each measured at several sizes, fitted
    t = a n^2 + b n^3 and EXTRAPOLATED to n (no CPU code can hold n = 50 000 in the reference's ~34 dense buffers).
      reference-literal: the unmodified reference classes (oracle/_ref/ref_driver; kind "reference") -- 3 dpotrf + 2 n x n dtrtrs
        = 3 n^3 flop through OpenBLAS plus ~100 element-wise n x n passes compiled -O0 like make_linux.  At the sample sizes the
        n^2 passes dominate, so the n^3 coefficient is not fitted from them: it is 3 flop / the OpenBLAS rate measured by the lean
        run in the same process (the reference cannot run its LAPACK calls faster than that);
      lean: the same mathematics with one dpotrf + one dpotri (oracle/lean_baseline.py): what a careful CPU implementation costs.
    `value` (and the reference arm's line) is the reference-literal extrapolation; `lean` is printed beside it."""
    from oracle import lean_baseline as LB
    threads = LB.blas_threads()
    LB.time_lean([1500])                                          # page OpenBLAS in
    lean = LB.time_lean(list(lean_sizes))
    gf = lean[-1][2]                                              # GFLOP/s of potrf + potri at the largest lean size
    a_l, b_l = LB.fit_n2_n3([r[0] for r in lean], [r[1] for r in lean])
    t_lean = a_l * n ** 2 + b_l * float(n) ** 3
    out = {"cores": cores, "blas_threads": threads, "extrapolated": True,
           "lean": {"what": "1 dpotrf + 1 dpotri + O(n^2) passes, scipy OpenBLAS, all cores (oracle/lean_baseline.py)",
                    "samples": [{"n": r[0], "seconds": round(r[1], 3), "gflops_potrf_potri": round(r[2], 1)} for r in lean],
                    "fit": {"a_n2": a_l, "b_n3": b_l}, "seconds_at_n": t_lean, "value": 1.0 / t_lean, "unit": "evals/s"}}
    if ref_samples:
        ns = [r[0] for r in ref_samples]
        ts = [r[1] for r in ref_samples]
        b_ref = 3.0 / (gf * 1e9)
        a_ref, b_ref = LB.fit_n2_n3(ns, ts, b_fixed=b_ref)
        t_ref = a_ref * n ** 2 + b_ref * float(n) ** 3
        out.update({"kind": "reference", "value": 1.0 / t_ref, "unit": "evals/s",
                    "sample": ("the unmodified reference (oracle/_ref/ref_driver: reference sources + Armadillo stand-in + OpenBLAS, built -O0 like "
                               "make_linux), set_GP_Pars + Grad_Values at n = %s: %s s per evaluation on %d host threads; fitted t = a n^2 + b n^3 with "
                               "b = 3 flop / the %.0f GFLOP/s OpenBLAS rate measured beside it, a = %.3g s by least squares; EXTRAPOLATED to n=%d: %.0f s "
                               "(a n^2 = %.0f s of -O0 element-wise passes, b n^3 = %.0f s of LAPACK); the reference itself cannot hold n=%d "
                               "(~34 dense n x n buffers = ~680 GB)"
                               % ("/".join(str(v) for v in ns), "/".join("%.2f" % v for v in ts), cores, gf, a_ref, n, t_ref,
                                  a_ref * n ** 2, b_ref * float(n) ** 3, n)),
                    "reference_literal": {"samples": [{"n": a, "seconds": round(b, 3)} for a, b in ref_samples],
                                          "fit": {"a_n2": a_ref, "b_n3": b_ref}, "seconds_at_n": t_ref}})
    else:
        out.update({"kind": "port", "value": 1.0 / t_lean, "unit": "evals/s",
                    "sample": "oracle/_ref/ref_driver not present: the lean numpy/LAPACK port only (see `lean`), extrapolated to n=%d" % n})
    return out


def ref_samples_once(sizes):
    """One timed evaluation of the compiled reference per size (after one untimed one at the smallest size)."""
    if not os.path.exists(REF_DRIVER):
        return []
    cpu_reference_times(sizes[0], 0, 1)
    out = []
    for ns in sizes:
        secs, _ = cpu_reference_times(ns, 0, 1)
        out.append((ns, float(np.mean(secs))))
    return out


def recorded_fits():
    """BASELINE configs[2] (LBFGS fit at n = 50 000 through the reference's command line) is minutes of GPU time per run, so the bench line carries
    the RECORDED runs of scripts/fit_n50k.py (profiles/r02_fit_n50k_<N>gpu.json) instead of repeating them: not measured in this run."""
    runs = []
    for g in (1, 2, 4, 8):
        path = os.path.join(ROOT, "profiles", "r02_fit_n50k_%dgpu.json" % g)
        if not os.path.exists(path):
            continue
        try:
            with open(path) as f:
                r = json.loads(f.readline())
        except (OSError, ValueError):
            continue
        runs.append({k: r.get(k) for k in ("gpus", "iters", "wall_s", "device_s", "objective_calls", "gradient_calls", "nlml_first", "nlml_last")})
    if not runs:
        return None
    return {"what": "gp_ss_ak -v 3 -pm 1 train -k ExpAns -kn 1 -o LBFGS -# ITERS on a synthetic n = 50 000 file, end to end (scripts/fit_n50k.py)",
            "recorded": True, "source": "profiles/r02_fit_n50k_<N>gpu.json", "runs": runs}


def run_reference_arm(args, rank, world):
    """bench.py --impl reference: the reference's own CPU implementation on the host cores.  Each of the W + K steps is ONE
    LML+gradient evaluation of the compiled reference at one of REF_SIZES (cycled); the K timed ones are averaged per size, fitted
    and extrapolated to n as cpu_baseline_block documents.  ms_per_step is the extrapolated time of one evaluation at n."""
    if rank != 0:
        return
    cores = os.cpu_count()
    sizes = REF_SIZES if os.path.exists(REF_DRIVER) else ()
    per = {ns: [] for ns in sizes}
    for k in range(args.warmup + args.steps):
        if not sizes:
            break
        ns = sizes[k % len(sizes)]
        secs, _ = cpu_reference_times(ns, 0, 1)
        if k >= args.warmup:
            per[ns].append(secs[0])
    samples = [(ns, float(np.mean(v))) for ns, v in per.items() if v]
    cb = cpu_baseline_block(args.n, samples, LEAN_SIZES, cores)
    value = cb["value"]
    line = {
        "impl": "reference", "metric": "ExpAns LML+grad evals/s at n=%dk" % (args.n // 1000), "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "LML+gradient evaluation (set_GP_Pars + Grad_Values), ExpAns+Bias 3-D, n=%d; each step is a bounded sample: one "
                               "evaluation of the compiled reference at n in %s (cycled), fitted a n^2 + b n^3 and extrapolated to n=%d"
                               % (args.n, list(sizes), args.n), "n": args.n, "n_samples": list(sizes)},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." from C when NCCL_DEBUG=VERSION is in the
# environment, as on the GPU boxes of round 2), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved one.
_REAL_STDOUT = None


def guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=N_HEADLINE)
    ap.add_argument("--impl", default="gpss")
    ap.add_argument("--mode", default="distributed", choices=["distributed", "replicas"],
                    help="N > 1: one evaluation partitioned over all GPUs (default), or one independent evaluation stream per GPU")
    ap.add_argument("--cpu-lean-max", type=int, default=0, help="largest extra size of the lean CPU baseline (e.g. 20000; default: 5000 and 10000 only)")
    ap.add_argument("--pred-m", type=int, default=32768, help="test points PER GPU of the prediction leg (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gp_ss_ak_b200 as G
    from gp_ss_ak_b200 import datagen

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    distributed = world > 1 and args.mode == "distributed"
    # distributed: every rank holds the SAME data and the same theta sequence (one evaluation spread over all GPUs);
    # replicas: every rank has its own data set and theta sequence.
    stream_id = 0 if (distributed or world == 1) else rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.n
    X, y = datagen.drillholes(n, seed=stream_id)
    Xs, ys, std_params = datagen.standardise_symmetric(X, y)
    Xf = np.asfortranarray(Xs)
    model = G.GpssModel(Xf, ys, device=local_rank)
    n_pad = model.padded_n()
    if distributed:
        ids = [G.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        model.dist_init(rank, world, ids[0])

    def checked_eval():
        """One Grad_Values; a NaN (Cholesky failure path) would skip trtri/lauum and fake the timing, so it is fatal."""
        L, g = model.nlml_grad()
        if not (np.isfinite(L) and np.all(np.isfinite(g))):
            raise SystemExit("bench.py: evaluation returned a non-finite objective/gradient (L=%r) -- refusing to time it" % L)
        return L, g

    # ---- warm-up (>= 3): also allocates U / Q and pages the kernels in ----
    for k in range(args.warmup):
        model.set_theta(theta_probe(stream_id * 7919 + k))
        checked_eval()

    # ---- timed region 1: device-resident (`value`) ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = model.launch_count()
    barrier()
    dev_ms = []
    t0 = time.perf_counter()
    for k in range(args.steps):
        model.set_theta(theta_probe(stream_id * 7919 + 100 + k))
        L, g = checked_eval()
        dev_ms.append(model.last_call_ms())      # CUDA events on the stream the kernels are launched on
    barrier()
    wall = time.perf_counter() - t0
    launches = model.launch_count() - l0
    t_dev = float(np.sum(dev_ms)) * 1e-3

    # ---- timed region 2: end to end through the C ABI with HOST buffers (`e2e`) ----
    barrier()
    t1 = time.perf_counter()
    for k in range(args.steps):
        model.set_data(Xf, ys)                    # host -> device copy of this step's inputs
        model.set_theta(theta_probe(stream_id * 7919 + 200 + k))
        L, g = checked_eval()                     # device -> host read of value + g[10]
    barrier()
    wall_e2e = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel (oz_gemm_kernel on the default int8 pipe, gemm_nt_ws_kernel with GPSS_OZAKI=0) from a profiled evaluation ----
    model.set_profiling(True)
    model.set_theta(theta_probe(stream_id * 7919 + 300))
    checked_eval()
    ph = model.phase_ms()
    model.set_profiling(False)
    gemm_ms = float(ph[1] + ph[3] + ph[4])        # potrf + trtri + lauum phases: >99% of it inside gemm_nt_ws_kernel
    share = world if distributed else 1
    alg_flops = float(n_pad) ** 3 / share         # n^3/3 each; per GPU when the evaluation is partitioned
    try:
        peak = G.measure_fp64_peak(local_rank)
        peak_src = "DMMA.8x8x4 register-loop micro-peak measured live by gpss_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 entry)"
    except Exception:
        peak = FP64_PEAK_FALLBACK_TFLOPS
        peak_src = "profiles/r01_fp64_peak_microbench.txt"
    achieved = alg_flops / (gemm_ms * 1e-3) * 1e-12
    oz_s = model.ozaki_slices()                   # 0: FP64 DMMA; 6 | 7 | 8: int8 tensor cores, that many 7-bit slices (gpss_ozaki.cuh)

    # ---- prediction leg (BASELINE metric iii): mean + variance of block-model centroids, test points split over the GPUs,
    #      L / alpha replicated; host buffers in, host buffers out (the public call), device time from the handle's events ----
    pred = None
    if args.pred_m > 0:
        m_total = args.pred_m * world
        side = int(round(m_total ** (1.0 / 3.0))) + 1
        grid = datagen.block_model(side, side, side, X.min(axis=0), X.max(axis=0))[:m_total]
        Xt = np.asfortranarray((grid - std_params[1:, 0]) / std_params[1:, 1])
        sums = np.array([math.fsum(Xt[:, j]) for j in range(3)])
        lo, hi = rank * args.pred_m, (rank + 1) * args.pred_m
        shard = np.asfortranarray(Xt[lo:hi])
        model.predict_shard(m_total, sums, shard[:8192])            # warm-up: builds W = L^-1, allocates the batch buffers
        barrier()
        tp0 = time.perf_counter()
        mu_s, var_s = model.predict_shard(m_total, sums, shard)
        t_pred_dev = model.last_call_ms() * 1e-3
        barrier()
        t_pred_wall = time.perf_counter() - tp0
        if not (np.all(np.isfinite(mu_s)) and np.all(np.isfinite(var_s))):
            raise SystemExit("bench.py: non-finite prediction")
        pred = [t_pred_dev, t_pred_wall]

    # ---- max over ranks ----
    phases_all = None
    if world > 1:
        phases_all = [None] * world
        dist.all_gather_object(phases_all, [round(float(v), 2) for v in (ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[8])])
        vals = [t_dev, wall, wall_e2e, gemm_ms] + (pred or [0.0, 0.0])
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, wall, wall_e2e, gemm_ms_max, p0, p1 = (float(v) for v in t.cpu())
        if pred:
            pred = [p0, p1]
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank == 0:
        total_evals = args.steps * (1 if distributed else world)
        value = total_evals / t_dev
        e2e_value = total_evals / wall_e2e
        if world == 1:
            par = "single GPU"
        elif distributed:
            par = ("one evaluation over %d GPUs: 512-wide block columns of the Cholesky factor owned round-robin, NCCL panel broadcast, "
                   "row-sliced triangular inverse / B^-1, all-reduced gradient partials" % world)
        else:
            par = "replicas x%d (independent evaluations)" % world
        note = ("achieved = n_pad^3%s algorithmic flops (potrf+trtri+lauum) / device time of those phases on rank 0"
                % (" / %d GPUs" % world if distributed else ""))
        if oz_s:
            pairs = oz_s * (oz_s + 1) // 2
            i8_peak, i8_burst, i8_src = int8_tensor_peak(local_rank)
            # executed int8 op/s: exact tile-level count on one GPU; a partitioned evaluation is approximated by its share of n^3
            i8_ops = 2.0 * int8_macs(n_pad) * pairs if share == 1 else alg_flops * pairs
            i8_tops = i8_ops / (gemm_ms * 1e-3) * 1e-12
            # the kernel alone: the lauum phase is oz_gemm_kernel launches and nothing else (one launch with 7-bit digits)
            lau_ops = 2.0 * int8_macs_lauum(n_pad) * pairs / share
            lau_tops = lau_ops / (float(ph[4]) * 1e-3) * 1e-12
            roofline = {"bound": "tensor", "achieved": lau_tops, "peak": i8_peak, "unit": "TOP/s (int8 tensor pipe)",
                        "frac": lau_tops / i8_peak, "traffic": OZ_LAUUM_TRAFFIC_BYTES if share == 1 else None,    # measured for the one-GPU launch only
                        "traffic_note": oz_traffic_note(oz_s, n_pad),
                        "kernel": "oz_gemm_kernel<%d, 64, merged> (tcgen05.mma kind::i8 M 128 N <= 256, TMA SWIZZLE_64B operand planes, int32 accumulators in TMEM; "
                                  "%d int8 products per FP64 product)" % (oz_s, pairs),
                        "launch": "B^-1 = U U^T (lauum phase: oz_gemm_kernel launches only), %.3e int8 op in %.1f ms, CUDA events on the handle's stream"
                                  % (lau_ops, float(ph[4])),
                        "peak_source": i8_src, "peak_burst": i8_burst, "frac_of_burst": (lau_tops / i8_burst) if i8_burst else None,
                        "all_gemm_phases": {"achieved": i8_tops, "frac": i8_tops / i8_peak,
                                            "note": "int8 ops of ALL oz_gemm_kernel launches of one evaluation (tile-level count, %.1f %% of n_pad^3, x %d slice pairs) / device "
                                                    "time of the potrf + trtri + lauum phases, which also hold the k = 512 panel work on the DMMA pipe and the digit slicing"
                                                    % (100.0 * i8_ops / pairs / alg_flops, pairs)},
                        "fp64_equivalent_tflops": achieved, "fp64_dmma_peak_tflops": peak, "frac_of_fp64_dmma_peak": achieved / peak,
                        "note": "achieved = int8 ops of the B^-1 = U U^T launch / its duration; fp64_equivalent_tflops = n_pad^3 / (potrf + trtri + lauum time)"}
            oz_bits = model.ozaki_digit_bits()
            dtype = "f64 (long-k products as %d x %d-bit int8 slices on the tensor cores, int32 accumulation, FP64 recombination)" % (oz_s, oz_bits)
            gemm_path = "int8 tensor cores, Ozaki splitting, %d slices of %d bits (GPSS_OZAKI / GPSS_OZAKI_BITS; GPSS_OZAKI=0: FP64 DMMA)" % (oz_s, oz_bits)
        else:
            roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                        "traffic": GEMM_TRAFFIC_BYTES, "traffic_note": GEMM_TRAFFIC_NOTE,
                        "kernel": "gemm_nt_ws_kernel<GemmTileWS<128,64,2,2,2,4>> (FP64 DMMA, bulk-copy producer warp + mbarrier ring)",
                        "peak_source": peak_src, "note": note}
            dtype = "f64"
            gemm_path = "FP64 DMMA"
        line = {
            "metric": "ExpAns LML+grad evals/s at n=%dk" % (n // 1000), "value": value, "unit": "evals/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if (distributed or world == 1) else "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic",
            "config": {"workload": "LML+gradient evaluation (set_GP_Pars + Grad_Values), ExpAns+Bias 3-D, n=%d, theta differs every step"
                                   % n, "n": n, "n_pad": n_pad, "l2_policy": "inputs larger than L2 (K, L, B^-1 = %.1f GB each)"
                                   % (n_pad * n_pad * 8 / 1e9), "parallelism": par, "gemm_path": gemm_path},
            "wall_ms_per_step": wall / args.steps * 1e3,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(n * 4 * 8 + 80) * (world if world > 1 else 1),
                    "d2h_bytes_per_step": 88 * (world if world > 1 else 1)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roofline,
            "phases_ms": {"kbuild": ph[0], "potrf": ph[1], "solve_objective": ph[2], "trtri": ph[3], "lauum": ph[4], "grad_pass": ph[5],
                          "gather_U": ph[8]},
            "cholesky_tflops": (float(n_pad) ** 3 / 3) / (ph[1] * 1e-3) * 1e-12,
        }
        if phases_all:
            line["phases_ms_per_rank"] = {"order": ["kbuild", "potrf", "solve_objective", "trtri", "lauum", "grad_pass", "gather_U"],
                                          "ranks": phases_all}
        if pred:
            m_total = args.pred_m * world
            flop_pt = float(n_pad) ** 2       # one triangular solve per point, n^2/2 FMA
            fp64_eq = m_total / world * flop_pt / pred[0] * 1e-12
            if oz_s and os.environ.get("GPSS_OZAKI_PREDICT", "1") != "0":
                pairs_p = oz_s * (oz_s + 1) // 2
                pred_roof = {"bound": "tensor", "achieved": fp64_eq * pairs_p, "peak": i8_peak, "unit": "TOP/s (int8 tensor pipe)",
                             "frac": fp64_eq * pairs_p / i8_peak, "fp64_equivalent_tflops": fp64_eq, "fp64_dmma_peak_tflops": peak,
                             "note": "n_pad^2 FP64-equivalent flop per test point (triangular k-range of W = L^-1) x %d slice pairs on oz_gemm_kernel, "
                                     "per GPU; the time also holds the cross-covariance build, its digit slicing and the variance reduction" % pairs_p}
            else:
                pred_roof = {"bound": "tensor", "achieved": fp64_eq, "peak": peak, "unit": "TFLOP/s", "frac": fp64_eq / peak,
                             "note": "n_pad^2 flop per test point (triangular k-range of W = L^-1), per GPU"}
            line["predict"] = {"metric": "test predictions/s (mean + variance) vs the n=%d model" % n, "value": m_total / pred[0],
                               "unit": "preds/s", "m_total": m_total, "m_per_gpu": args.pred_m, "device_ms": pred[0] * 1e3,
                               "e2e": {"value": m_total / pred[1], "unit": "preds/s", "h2d_bytes": 24 * m_total, "d2h_bytes": 16 * m_total},
                               "roofline": pred_roof,
                               "sharding": "test points split over %d GPUs, L / alpha replicated" % world}
        fit = recorded_fits()
        if fit:
            line["fit"] = fit
        line["phases_note"] = ("phases_ms come from one PROFILED evaluation after the timed region: phase timers synchronise the host after every phase, "
                               "so the solves do not overlap the inverse there (and, on multi-GPU handles, the inverse is not issued inside the Cholesky); "
                               "their sum exceeds ms_per_step by what those overlaps save")
        if not args.no_cpu_baseline and world == 1:      # the CPU legs run on rank 0 at N = 1 only (other ranks would sit in NCCL teardown meanwhile)
            lean_sizes = LEAN_SIZES + ((args.cpu_lean_max,) if args.cpu_lean_max > LEAN_SIZES[-1] else ())
            line["cpu_baseline"] = cpu_baseline_block(n, ref_samples_once(REF_SIZES), lean_sizes, os.cpu_count())
        emit(line)
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
