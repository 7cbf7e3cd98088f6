// gpss_potrf.cuh -- part of the single translation unit gpss_capi.cu: the blocked FP64 Cholesky drivers (single GPU with
// look-ahead, distributed with replicated factor, partitioned storage) and the partitioned triangular solves.
#pragma once
// ---------------------------------------------------------------------------------------------------
// blocked right-looking Cholesky, two-level (outer NBO = 512 for deep-k trailing updates, inner 128)
// A: n_pad x n_pad lower, in place.  Replaces arma::chol -> dpotrf (GP_Utils.cpp:881,903).
// ---------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------
// distributed evaluation helpers (world > 1): staging buffer, balanced row partitions
// ---------------------------------------------------------------------------------------------------
static int ensure_stage(gpss_ctx* c, size_t count)
{
  if (c->stage_count >= count) return GPSS_OK;
  if (c->stage) cudaFree(c->stage);
  c->stage = nullptr;
  c->stage_count = 0;
  CU(cudaMalloc(&c->stage, count * sizeof(double)));
  c->stage_count = count;
  return GPSS_OK;
}

// Row boundaries (multiples of 128) that give every rank the same share of work:
//   kind 0: rows of U = L^-T in the block-column inverse, cost(row i) ~ (n - i)^2 / 2   -> (n - r_k)^3 = n^3 (1 - k/P)
//   kind 1: rows of B^-1 = U U^T (lower),                 cost(row i) ~ (i + 1)(n - i)  -> n x^2/2 - x^3/3 = (k/P) n^3/6
//   kind 2: rows of U = L^-T when the bulk product of the inverse runs on oz_gemm_kernel (ONE CTA per SM, 128 x 64 tiles, NBO / 64 = 8 tile
//           columns per step).  A slice of t tile rows is 8 t CTAs per step: up to 18 tile rows it is one partial wave whose length is its
//           LONGEST k-range, 19 tile rows are two waves -- the flop-balanced partition of kind 0 gives ranks 1-4 of 8 exactly such slices at
//           n = 50 000 (measured: 94 / 85 / 78 / 69 ms against 58 ms for rank 0, profiles/r02_dist_8gpu.log).  Model of a step at block column
//           J0 for rows [a, a + t): its CTAs run in waves of 148, each wave as long as its first (longest) k-range, summed over the steps;
//           the boundaries minimise the largest rank cost (bisection on the cost, greedy assignment).
static double inv_slice_cost(int a_t, int t, int ntile)
{
  const int tpb = NBO / NB;                                  // tile rows per block column
  const int tc = NBO / 64;                                   // tile columns per step
  double cost = 0.0;
  for (int J0t = tpb; J0t < ntile; J0t += tpb) {             // block columns 1 .. : rows above the block are [0, J0t)
    const int r1 = (a_t + t < J0t) ? a_t + t : J0t;
    if (r1 <= a_t) continue;
    // CTAs are scheduled in row order, longest k-range first (row a_t), 148 at a time: wave w lasts as long as its first CTA, which
    // belongs to tile row a_t + floor(148 w / tc)
    const int jobs = (r1 - a_t) * tc;
    for (int w = 0; w * 148 < jobs; w++) cost += J0t - (a_t + (w * 148) / tc);
  }
  return cost;
}

static void balanced_rows(int n_pad, int world, int kind, std::vector<int>& bounds)
{
  bounds.assign(world + 1, 0);
  bounds[world] = n_pad;
  if (kind == 2) {
    const int ntile = n_pad / NB;
    double lo = 0.0, hi = inv_slice_cost(0, ntile, ntile);
    std::vector<int> best(world + 1, 0);
    best[world] = ntile;
    for (int k = 1; k < world; k++) best[k] = (int)((long)ntile * k / world);
    for (int it = 0; it < 60; it++) {
      const double C = 0.5 * (lo + hi);
      std::vector<int> b(world + 1, 0);
      int a = 0;
      for (int k = 0; k + 1 < world; k++) {                  // the most tile rows whose cost stays within C (the cost grows with t)
        int t_lo = 0, t_hi = ntile - a;
        while (t_lo < t_hi) {
          const int mid = (t_lo + t_hi + 1) / 2;
          if (inv_slice_cost(a, mid, ntile) <= C) t_lo = mid; else t_hi = mid - 1;
        }
        a += t_lo;
        b[k + 1] = a;
      }
      b[world] = ntile;
      if (inv_slice_cost(a, ntile - a, ntile) <= C) { hi = C; best = b; } else lo = C;
    }
    for (int k = 0; k <= world; k++) bounds[k] = best[k] * NB;
    return;
  }
  const double n = n_pad;
  for (int k = 1; k < world; k++) {
    const double f = (double)k / world;
    double x;
    if (kind == 0) {
      x = n * (1.0 - std::cbrt(1.0 - f));
    } else {
      double lo = 0.0, hi = n;
      const double target = f * n * n * n / 6.0;
      for (int it = 0; it < 100; it++) {
        const double mid = 0.5 * (lo + hi);
        if (n * mid * mid / 2.0 - mid * mid * mid / 3.0 < target) lo = mid; else hi = mid;
      }
      x = 0.5 * (lo + hi);
    }
    int b = (int)std::lround(x / NB) * NB;
    if (b < bounds[k - 1]) b = bounds[k - 1];
    if (b > n_pad) b = n_pad;
    bounds[k] = b;
  }
}

// ---------------------------------------------------------------------------------------------------
// The distributed Cholesky as a per-rank list of operations (pure host logic; gpss_dist_potrf_schedule exposes it so the
// CPU tests can replay all ranks and check that every block column sees every earlier panel exactly once, in an order
// the broadcasts make possible).  Block column j (owner j % P) receives, all on ONE low-priority side stream (they
// update the same tiles, so they serialise anyway):
//     chunk A(j):   panels 0 .. j-P        one long-k GEMM, issued as soon as the owner has factored its previous column
//     single(j,t):  panel t, j-P < t < j-1 (k = NBO), issued when panel t arrives
// and on the main stream U2(j) = panel j-1, the panel factorisation and the broadcast.  A rank therefore always has
// about P panel periods of bulk work queued behind the critical path instead of one.
// ---------------------------------------------------------------------------------------------------
enum { DIST_WAIT_SIDE = 0, DIST_UPDATE_MAIN = 1, DIST_FACTOR = 2, DIST_BCAST = 3, DIST_UPDATE_SIDE = 4 };
struct DistOp { int kind, col, pbeg, pcnt, root, stream; };   // update ops apply panels pbeg .. pbeg+pcnt-1 to block column col
static void dist_potrf_schedule(int nblk_o, int P, int me, std::vector<DistOp>& ops)
{
  ops.clear();
  for (int t = 0; t < nblk_o; t++) {
    const bool mine = (t % P) == me;
    if (mine) {
      if (t >= 2) ops.push_back({DIST_WAIT_SIDE, t, 0, 0, 0, 0});
      if (t >= 1) ops.push_back({DIST_UPDATE_MAIN, t, P == 1 ? 0 : t - 1, P == 1 ? t : 1, 0, 0});   // alone: plain left-looking
      ops.push_back({DIST_FACTOR, t, 0, 0, 0, 0});
    }
    ops.push_back({DIST_BCAST, t, 0, 0, t % P, 0});
    int j = t + ((me - t) % P + P) % P;        // my next block column after t
    if (j == t) j = t + P;
    if (j >= nblk_o || t >= j - 1) continue;   // panel j-1 is U2(j)
    const int stream = (j / P) & 1;
    if (mine) ops.push_back({DIST_UPDATE_SIDE, j, 0, t + 1, 0, stream});     // chunk A
    else ops.push_back({DIST_UPDATE_SIDE, j, t, 1, 0, stream});              // one panel
  }
}

// One outer panel: factor the NBO-wide block column starting at K0 (all rows below), 128 columns at a time.
template <class StepDone>
static int potrf_panel(gpss_ctx* c, double* A, long ld, int n_pad, int K0, int nbk, double* Winv, double* logdet_parts, int* dflag,
                       StepDone step_done)
{
  for (int k = K0; k < K0 + nbk; k += NB) {
    double* Akk = A + (long)k * ld + k;
    double* Wk = Winv + (long)(k / NB) * NB * NB;
    potrf_diag_inv_kernel<<<1, DIAG_THREADS, DIAG_SMEM, c->st>>>(Akk, ld, Wk, logdet_parts + k / NB, dflag);
    c->launches++;
    CU(cudaGetLastError());
    const int m = n_pad - k - NB;
    if (m <= 0) { RET(step_done(k)); continue; }
    double* A21 = A + (long)k * ld + (k + NB);
    {  // panel solve, in place: A21 <- A21 * inv(L11)^T, with the 128x64 tile (it shares an SM with a resident
       // trailing-update CTA, which the 128x128 tile cannot).  Columns 64..127 first: they read all 128 input
       // columns; columns 0..63 then need only inputs 0..63 because inv(L11) is lower triangular.
      GemmArgs g1 = gemm_args(A21, ld, Wk + 64, NB, A21 + 64 * ld, ld, m, 64, NB);
      RET(gemm(c, g1));
      GemmArgs g2 = gemm_args(A21, ld, Wk, NB, A21, ld, m, 64, 64);
      RET(gemm(c, g2));
    }
    RET(step_done(k));            // columns k .. k+127 of the factor are final from here on
    const int ncols = K0 + nbk - (k + NB);
    if (ncols > 0) {  // update of the remaining columns of the outer panel
      double* A22 = A + (long)(k + NB) * ld + (k + NB);
      GemmArgs g = gemm_args(A21, ld, A21, ld, A22, ld, m, ncols, NB);
      g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = k + NB; g.gcol0 = k + NB;
      RET(gemm(c, g));
    }
  }
  return GPSS_OK;
}

static int potrf_panel(gpss_ctx* c, double* A, long ld, int n_pad, int K0, int nbk, double* Winv, double* logdet_parts, int* dflag)
{
  return potrf_panel(c, A, ld, n_pad, K0, nbk, Winv, logdet_parts, dflag, [](int) { return (int)GPSS_OK; });
}

// ---------------------------------------------------------------------------------------------------
// Distributed critical path, fused form.  A block column is factored in two parts with different latencies:
//   (1) potrf_diag_chain: ONLY the nbk x nbk diagonal block D -- the 128-steps of potrf_panel with every GEMM cut down to the rows of
//       D (<= 384), then W = inv(L_D) as one nbk x nbk lower-triangular matrix (trtri_diag_block + a transpose).  ~30 dependent launches
//       on <= 2 MB of data: pure latency, independent of n.
//   (2) ONE full-height GEMM  A[below D, block column] <- A[below D, block column] W^T  (k = nbk, triangular k-range), written
//       out of place -- straight into the broadcast buffer, so the pack kernel disappears too.
// The old panel made the full height wait for each of the four diagonal steps (16 dependent full-height launches, 0.73 ms at n = 50 000);
// here the full-height work is one launch after the chain, and the owner applies the full-height part of U2 on a second stream WHILE the
// chain runs (potrf_blocked).  Same flops (m x 512^2 / 2 x 2 against m x 163 840 x 2).
// ---------------------------------------------------------------------------------------------------
static int trtri_diag_block(gpss_ctx* c, double* U, long ldu, const double* L, long ldl, int J0, int nbj);
static int trtri_step(gpss_ctx* c, const TrtriRun& R, int t);
static int potrf_diag_chain(gpss_ctx* c, double* A, long ld, int K0, int nbk, double* Winv, double* logdet_parts, int* dflag)
{
  for (int k = K0; k < K0 + nbk; k += NB) {
    double* Akk = A + (long)k * ld + k;
    double* Wk = Winv + (long)(k / NB) * NB * NB;
    potrf_diag_inv_kernel<<<1, DIAG_THREADS, DIAG_SMEM, c->st>>>(Akk, ld, Wk, logdet_parts + k / NB, dflag);
    c->launches++;
    CU(cudaGetLastError());
    const int m = K0 + nbk - k - NB;                    // rows of D below this step
    if (m <= 0) continue;
    double* A21 = A + (long)k * ld + (k + NB);
    GemmArgs g1 = gemm_args(A21, ld, Wk + 64, NB, A21 + 64 * ld, ld, m, 64, NB);
    RET(gemm(c, g1));
    GemmArgs g2 = gemm_args(A21, ld, Wk, NB, A21, ld, m, 64, 64);
    RET(gemm(c, g2));
    double* A22 = A + (long)(k + NB) * ld + (k + NB);
    GemmArgs g = gemm_args(A21, ld, A21, ld, A22, ld, m, m, NB);
    g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = k + NB; g.gcol0 = k + NB;
    RET(gemm(c, g));
  }
  // W = inv(L_D): the transposed inverse in 128-steps (as the diagonal blocks of U = L^-T), then transposed into column-major lower form
  double* Ud = c->Wpan;
  double* Wd = c->Wpan + (size_t)NBO * NBO;
  RET(trtri_diag_block(c, Ud - ((long)K0 * NBO + K0), NBO, A, ld, K0, nbk));
  transpose_kernel<<<dim3(nbk / 32, nbk / 32), 256, 0, c->st>>>(Wd, NBO, Ud, NBO, 1);
  c->launches++;
  CU(cudaGetLastError());
  return GPSS_OK;
}

// LEFT-looking blocked Cholesky with look-ahead.  Block column T (width NBO) receives
//     U1(T):  A[T:, T] -= L[T:, 0:T-1] L[T, 0:T-1]^T     (panels 0..T-2: one long-k DMMA GEMM, side stream)
//     U2(T):  A[T:, T] -= L[T:, T-1]   L[T, T-1]^T       (panel T-1, k = NBO, main stream)
// and is then factored by potrf_panel on the main stream.  U1(T+1) only needs panels <= T-1, so it runs on the
// side stream WHILE the main stream does U2(T) and the latency-bound panel T: the DMMA pipe never waits for a
// panel, every output tile is written once per update instead of once per outer step (the right-looking form
// re-read and re-wrote the whole trailing matrix n/NBO times), and nearly all flops run in long-k GEMMs.
static int potrf_blocked(gpss_ctx* c, double* A, long ld, int n_pad, double* Winv, double* logdet_parts, int* dflag)
{
  const int P = c->world, me = c->rank;
  const bool la = c->st2 != nullptr && !getenv("GPSS_NO_LOOKAHEAD");
  if (P > 1 && !la) return fail_arg("the distributed factorisation needs the look-ahead streams");
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  if (la && (int)c->ev_pool.size() < 2 * nblk_o + 2) {
    const size_t want = 2 * nblk_o + 2;
    while (c->ev_pool.size() < want) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->ev_pool.push_back(e);
    }
  }
  // staging layout of one broadcast: [panel rows T0.. x nbT | the panel's 128x128 diagonal inverses | their log-dets]
  const size_t stage_need = (size_t)n_pad * NBO + (size_t)(NBO / NB) * NB * NB + NBO / NB;
  if (P > 1) RET(ensure_stage(c, stage_need));
  // int8 path (gpss_ozaki.cuh): every finished block column is cut into digit planes on the main stream, and the long-k
  // look-ahead update U1 reads those planes through the tcgen05 kernel; U2 (k = NBO, critical path) and the panel stay on DMMA
  const bool pipe_env = getenv("GPSS_DIST_PIPE") != nullptr && atoi(getenv("GPSS_DIST_PIPE")) != 0;
  const bool ozk = oz_active(c) && c->ozL && la && A == c->Lm && n_pad == c->n_pad && ld == (long)c->n_pad && !(P > 1 && pipe_env);
  if (A == c->Lm) c->ozL_valid = ozk;                          // every panel is cut below iff ozk; the inverse must not read stale planes
  // main-stream DMMA GEMMs next to resident int8 CTAs (gemm_ws_on): only the 7-plane stage ring leaves the room
  struct Coresident {
    gpss_ctx* c;
    Coresident(gpss_ctx* c_, bool on) : c(c_) { c->dmma_coresident = on; }
    ~Coresident() { c->dmma_coresident = false; }
  } coresident(c, ozk && c->oz_s <= 7 && !(getenv("GPSS_DMMA_CORESIDENT") && atoi(getenv("GPSS_DMMA_CORESIDENT")) == 0));
  auto update = [&](int T0, int nbT, int kbeg, int klen, cudaStream_t stream) -> int {
    // A[T0:, T0:T0+nbT] -= L[T0:, kbeg:kbeg+klen] L[T0:T0+nbT, kbeg:kbeg+klen]^T
    // (distributed: only the long chunks -- a single received panel, k = NBO, stays on DMMA)
    // U2 (k = NBO, main stream) on the int8 kernel as well (GPSS_OZ_U2=0: DMMA) -- panel t-1 was cut into planes right after it was factored / received
    static const bool oz_u2 = !(getenv("GPSS_OZ_U2") != nullptr && atoi(getenv("GPSS_OZ_U2")) == 0);   // default on (454 vs 468 ms potrf at n = 50 000)
    if (ozk && ((stream != c->st && (P == 1 || klen >= 4 * NBO)) || (oz_u2 && stream == c->st))) {
      oz::Args a;
      memset(&a, 0, sizeof a);
      a.C = A + (long)T0 * ld + T0; a.ldc = ld; a.m = n_pad - T0; a.n = nbT;
      a.a_row0 = T0; a.b_row0 = T0; a.k0 = kbeg; a.k1 = kbeg + klen;
      a.lower_only = 1; a.grow0 = T0; a.gcol0 = T0;
      a.accumulate = 1; a.sign = -1.0; a.a_kind = oz::SCALE_CHOL; a.b_kind = oz::SCALE_CHOL;
      return oz_gemm_on(c, c->oz_tmL[0], c->oz_tmL[1], a, stream);
    }
    const double* Lp = A + (long)kbeg * ld + T0;
    GemmArgs g = gemm_args(Lp, ld, Lp, ld, A + (long)T0 * ld + T0, ld, n_pad - T0, nbT, klen);
    g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = T0; g.gcol0 = T0;
    return gemm_ws_on(c, g, stream);
  };
  // fused distributed panel: OPT-IN (GPSS_DIST_PANEL=fused).  Measured on 2 GPUs at n = 50 000 (profiles/r02_dist_fused_panel_2gpu.log): correct
  // (same results as one GPU to 1e-14 / 1e-12), but the chain is ~27 dependent launches of 10-17 us GEMMs + 4 x the diagonal kernel =
  // 0.94 ms per panel against 0.72 ms for the whole 16-launch panel, so an evaluation takes 759-769 ms against 744 ms.
  const bool fused = P > 1 && !pipe_env && Winv == c->Winv && getenv("GPSS_DIST_PANEL") && !strcmp(getenv("GPSS_DIST_PANEL"), "fused");
  // GPSS_DIST_U2=int8: the full-height part of U2 on the int8 kernel (its CTAs need whole SMs, which the bulk updates hold) instead of the
  // co-resident DMMA kernel
  const bool u2_int8 = fused && ozk && getenv("GPSS_DIST_U2") && !strcmp(getenv("GPSS_DIST_U2"), "int8");
  if (fused) {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    if (!c->st5) CU(cudaStreamCreateWithPriority(&c->st5, cudaStreamNonBlocking, hi));
    if (!c->st6) CU(cudaStreamCreateWithPriority(&c->st6, cudaStreamNonBlocking, hi));
    if (!c->ev_u2) CU(cudaEventCreateWithFlags(&c->ev_u2, cudaEventDisableTiming));
    if (!c->ev_unpacked) CU(cudaEventCreateWithFlags(&c->ev_unpacked, cudaEventDisableTiming));
    if (!c->Wpan) CU(cudaMalloc(&c->Wpan, sizeof(double) * 2 * NBO * NBO));
  }
  const bool fused1 = P == 1 && la && Winv == c->Winv && getenv("GPSS_PANEL") && !strcmp(getenv("GPSS_PANEL"), "fused");
  if (fused1) {
    if (!c->Wpan) CU(cudaMalloc(&c->Wpan, sizeof(double) * 2 * NBO * NBO));
    RET(ensure_stage(c, (size_t)n_pad * NBO));
  }
  // The inverse INSIDE the distributed factorisation.  With P ranks the tensor pipe of a rank has 1/P of a panel period's bulk work but waits
  // a whole period for the next panel (8 GPUs, n = 50 000: ~0.5 ms of updates per ~1.5 ms period).  Block column t of U = L^-T needs only
  // panels <= t, and a rank's rows of it need nothing from other ranks, so step t of the inverse is issued on two streams of its own (st8:
  // diagonal blocks, st9: bulk, lowest priority) the moment panel t is complete here.  Asked for by the overlapped gpss_nlml_grad
  // (want_trtri_interleaved; the buffers exist then); trtri_upper later only joins the streams.  GPSS_TRTRI_INTERLEAVE=0 / 1 forces it off / on.
  // Measured (profiles/r02_dist_8gpu.log, r02_dist_4gpu.log, r02_dist_interleaved_inverse_2gpu.log): 327.5 -> 304 ms per evaluation at 8 GPUs,
  // 436 -> 440 ms at 4, 731 -> 728 ms at 2 -- with few ranks the pipe has little idle time and the resident bulk CTAs of the inverse delay the
  // panel kernels (the diagonal-block kernel cannot share an SM with an int8 CTA) by as much as they hide.  Default: on from 6 ranks up.
  const char* inter_env = getenv("GPSS_TRTRI_INTERLEAVE");
  const bool inter_wanted = inter_env ? atoi(inter_env) != 0 : P >= 6;
  const bool inter = P > 1 && ozk && c->want_trtri_interleaved && c->ozU && c->Um && c->Tpanel && c->Wjj && A == c->Lm && !pipe_env && inter_wanted;
  TrtriRun inv_run = {nullptr, nullptr, &c->ev_pipe, true, nullptr, nullptr};
  if (inter) {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    if (!c->st8) CU(cudaStreamCreateWithPriority(&c->st8, cudaStreamNonBlocking, hi));
    if (!c->st9) CU(cudaStreamCreateWithPriority(&c->st9, cudaStreamNonBlocking, lo));
    while ((int)c->ev_pipe.size() < 2 * nblk_o + 2) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      c->ev_pipe.push_back(e);
    }
    inv_run.sm = c->st8; inv_run.ss = c->st9;
    c->ozU_valid = true;
  }
  auto inverse_step = [&](int col) -> int {                  // panel `col` is complete on this rank (ev_pool[2 col])
    if (!inter) return GPSS_OK;
    CU(cudaStreamWaitEvent(c->st8, c->ev_pool[2 * col], 0));
    return trtri_step(c, inv_run, col);
  };
  // rows [T0 + r0, T0 + r0 + rows) of block column T0 -= L[same rows, kbeg : kbeg + klen] L[T0 : T0 + nbT, kbeg : kbeg + klen]^T  (DMMA)
  auto update_rows = [&](int T0, int nbT, int r0, int rows, int kbeg, int klen, cudaStream_t stream) -> int {
    if (rows <= 0) return GPSS_OK;
    GemmArgs g = gemm_args(A + (long)kbeg * ld + T0 + r0, ld, A + (long)kbeg * ld + T0, ld, A + (long)T0 * ld + T0 + r0, ld, rows, nbT, klen);
    g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = T0 + r0; g.gcol0 = T0;
    return gemm_ws_on(c, g, stream);
  };
  if (P > 1) {
    std::vector<DistOp> ops;
    dist_potrf_schedule(nblk_o, P, me, ops);
    int kchunk = 1 << 30;
    if (const char* e = getenv("GPSS_DIST_KCHUNK")) { const int v = atoi(e); if (v >= NBO) kchunk = (v / NBO) * NBO; }
    // GPSS_DIST_TRACE: timing events around every main-stream step, summed per kind after the factorisation (diagnostic)
    const bool trace = getenv("GPSS_DIST_TRACE") != nullptr;
    std::vector<std::pair<int, cudaEvent_t>> marks;
    auto mark = [&](int what) {
      if (!trace) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, c->st);
      marks.push_back({what, e});
    };
    mark(-1);
    // GPSS_DIST_PIPE=1 (experimental): the panel travels in its four 128-column sub-panels, each broadcast -- on a separate
    // communication stream -- as soon as the owner's step has finalised it, and the next owner applies U2 in four k = 128
    // pieces as they arrive: U2 and 3/4 of the broadcast overlap the factorisation instead of following it.
    const bool pipe = getenv("GPSS_DIST_PIPE") != nullptr && atoi(getenv("GPSS_DIST_PIPE")) != 0;
    const int nsub_all = n_pad / NB;
    if (pipe) {
      if (!c->st4) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->st4, cudaStreamNonBlocking, hi));
      }
      while ((int)c->ev_pipe.size() < 2 * nsub_all + 2) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_pipe.push_back(e);
      }
      CU(cudaEventRecord(c->ev_main, c->st));                             // the K build precedes everything on the comm stream too
      CU(cudaStreamWaitEvent(c->st4, c->ev_main, 0));
    }
    for (const DistOp& op : ops) {
      const int T0 = op.col * NBO;
      const int nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
      if (pipe && (op.kind == DIST_UPDATE_MAIN || op.kind == DIST_FACTOR || op.kind == DIST_BCAST)) {
        if (op.kind == DIST_UPDATE_MAIN) {
          // U2 in k = 128 pieces, each as soon as its sub-panel has been received (op.pcnt == 1 whenever P > 1)
          const int Kp = op.pbeg * NBO;
          const int nbK = (n_pad - Kp < NBO) ? (n_pad - Kp) : NBO;
          for (int k0 = 0; k0 < nbK; k0 += NB) {
            CU(cudaStreamWaitEvent(c->st, c->ev_pipe[2 * ((Kp + k0) / NB) + 1], 0));
            RET(update(T0, nbT, Kp + k0, NB, c->st));
          }
        } else if (op.kind == DIST_FACTOR) {
          RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag, [&](int k) -> int {
            CU(cudaEventRecord(c->ev_pipe[2 * (k / NB)], c->st));           // sub-panel k is final on the owner
            return GPSS_OK;
          }));
        } else {
          const bool mine = op.root == me;
          for (int k = T0; k < T0 + nbT; k += NB) {
            const long rows = n_pad - k;
            const size_t n_panel = (size_t)rows * NB, n_w = (size_t)NB * NB;
            double* Wk = Winv + (size_t)(k / NB) * NB * NB;
            if (mine) {
              CU(cudaStreamWaitEvent(c->st4, c->ev_pipe[2 * (k / NB)], 0));
              pack_kernel<<<592, 256, 0, c->st4>>>(c->stage, A + (long)k * ld + k, ld, rows, NB);
              CU(cudaMemcpyAsync(c->stage + n_panel, Wk, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              CU(cudaMemcpyAsync(c->stage + n_panel + n_w, logdet_parts + k / NB, sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              c->launches++;
            }
            NC(g_nccl.Broadcast(c->stage, c->stage, n_panel + n_w + 1, ncclDouble, op.root, c->comm, c->st4));
            if (!mine) {
              unpack_kernel<<<592, 256, 0, c->st4>>>(A + (long)k * ld + k, ld, c->stage, rows, NB);
              CU(cudaMemcpyAsync(Wk, c->stage + n_panel, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              CU(cudaMemcpyAsync(logdet_parts + k / NB, c->stage + n_panel + n_w, sizeof(double), cudaMemcpyDeviceToDevice, c->st4));
              c->launches++;
            }
            CU(cudaEventRecord(c->ev_pipe[2 * (k / NB) + 1], c->st4));      // sub-panel k is complete on this rank
          }
          CU(cudaEventRecord(c->ev_pool[2 * op.col], c->st4));              // panel op.col is complete on this rank
        }
        continue;
      }
      if (fused && op.kind != DIST_WAIT_SIDE && op.kind != DIST_UPDATE_SIDE) {
        const long rows = n_pad - T0;
        const size_t n_panel = (size_t)rows * nbT, n_w = (size_t)(nbT / NB) * NB * NB, n_l = nbT / NB;
        double* Wt = Winv + (size_t)(T0 / NB) * NB * NB;
        if (op.kind == DIST_UPDATE_MAIN) {
          // U2 = panel T-1 (unpacked on this stream just before).  Its rows of the diagonal block first, on the main stream, so that the
          // latency chain can start; everything below on st5, concurrently with the chain.
          const int kb = op.pbeg * NBO, kl = op.pcnt * NBO;
          if (u2_int8) {
            RET(update(T0, nbT, kb, kl, c->st));                           // needs the planes of panel T-1: cut on this stream in DIST_BCAST
          } else {
            CU(cudaEventRecord(c->ev_main, c->st));
            CU(cudaStreamWaitEvent(c->st5, c->ev_main, 0));
            RET(update_rows(T0, nbT, 0, nbT, kb, kl, c->st));
            RET(update_rows(T0, nbT, nbT, (int)rows - nbT, kb, kl, c->st5));
            CU(cudaEventRecord(c->ev_u2, c->st5));
          }
          mark(1);
        } else if (op.kind == DIST_FACTOR) {
          RET(potrf_diag_chain(c, A, ld, T0, nbT, Winv, logdet_parts, dflag));
          mark(2);
          if (!u2_int8 && op.col >= 1) CU(cudaStreamWaitEvent(c->st, c->ev_u2, 0));
          // the solved rows below D go straight into the broadcast buffer: stage(r, j) = sum_{k <= j} A(T0 + nbT + r, T0 + k) W(j, k)
          if (rows > nbT) {
            GemmArgs g = gemm_args(A + (long)T0 * ld + T0 + nbT, ld, c->Wpan + (size_t)NBO * NBO, NBO, c->stage + nbT, rows, (int)rows - nbT, nbT, nbT);
            g.kend_col = 1;
            RET(gemm_ws_on(c, g, c->st));
          }
          copy2d_kernel<<<64, 256, 0, c->st>>>(c->stage, rows, A + (long)T0 * ld + T0, ld, nbT, nbT);
          CU(cudaMemcpyAsync(c->stage + n_panel, Wt, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
          CU(cudaMemcpyAsync(c->stage + n_panel + n_w, logdet_parts + T0 / NB, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
          c->launches++;
          mark(3);
        } else {   // DIST_BCAST
          const bool mine = op.root == me;
          NC(g_nccl.Broadcast(c->stage, c->stage, n_panel + n_w + n_l, ncclDouble, op.root, c->comm, c->st));
          mark(mine ? 4 : 5);
          if (mine) {                                                      // D is in place already; the rows below it come back from the buffer
            if (rows > nbT) {
              copy2d_kernel<<<592, 256, 0, c->st>>>(A + (long)T0 * ld + T0 + nbT, ld, c->stage + nbT, rows, rows - nbT, nbT);
              c->launches++;
            }
          } else {
            unpack_kernel<<<592, 256, 0, c->st>>>(A + (long)T0 * ld + T0, ld, c->stage, rows, nbT);
            CU(cudaMemcpyAsync(Wt, c->stage + n_panel, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            CU(cudaMemcpyAsync(logdet_parts + T0 / NB, c->stage + n_panel + n_w, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            c->launches++;
          }
          mark(6);
          if (ozk && !u2_int8) {
            // digit planes of the panel: read by the bulk updates only, so they are cut beside the critical path
            CU(cudaEventRecord(c->ev_unpacked, c->st));
            CU(cudaStreamWaitEvent(c->st6, c->ev_unpacked, 0));
            RET(oz_slice_on(c, A, ld, T0, n_pad - T0, T0, nbT, oz::SCALE_CHOL, oz::MASK_LOWER, c->ozL, c->st6));
            CU(cudaEventRecord(c->ev_pool[2 * op.col], c->st6));
          } else {
            if (ozk) RET(oz_slice_on(c, A, ld, T0, n_pad - T0, T0, nbT, oz::SCALE_CHOL, oz::MASK_LOWER, c->ozL, c->st));
            CU(cudaEventRecord(c->ev_pool[2 * op.col], c->st));
          }
          RET(inverse_step(op.col));
        }
        continue;
      }
      switch (op.kind) {
        case DIST_WAIT_SIDE:                                               // every side-stream update of my column
          CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * op.col + 1], 0));
          mark(0);
          break;
        case DIST_UPDATE_MAIN:                                             // U2: the panel just received, on the critical path
          RET(update(T0, nbT, op.pbeg * NBO, op.pcnt * NBO, c->st));
          mark(1);
          break;
        case DIST_FACTOR:
          RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag));
          mark(2);
          break;
        case DIST_BCAST: {
          // the owner's finished block column (+ its diagonal inverses and log-dets) goes to every rank: after the loop L,
          // Winv and logdet_parts are replicated.  One NCCL broadcast per panel (<= 205 MB at n = 50k), on the main stream.
          const bool mine = op.root == me;
          const long rows = n_pad - T0;
          const size_t n_panel = (size_t)rows * nbT, n_w = (size_t)(nbT / NB) * NB * NB, n_l = nbT / NB;
          double* Wt = Winv + (size_t)(T0 / NB) * NB * NB;
          if (mine) {
            pack_kernel<<<592, 256, 0, c->st>>>(c->stage, A + (long)T0 * ld + T0, ld, rows, nbT);
            CU(cudaMemcpyAsync(c->stage + n_panel, Wt, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            CU(cudaMemcpyAsync(c->stage + n_panel + n_w, logdet_parts + T0 / NB, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            c->launches++;
          }
          if (mine) mark(3);
          NC(g_nccl.Broadcast(c->stage, c->stage, n_panel + n_w + n_l, ncclDouble, op.root, c->comm, c->st));
          mark(mine ? 4 : 5);
          if (!mine) {
            unpack_kernel<<<592, 256, 0, c->st>>>(A + (long)T0 * ld + T0, ld, c->stage, rows, nbT);
            CU(cudaMemcpyAsync(Wt, c->stage + n_panel, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            CU(cudaMemcpyAsync(logdet_parts + T0 / NB, c->stage + n_panel + n_w, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
            c->launches++;
          }
          if (ozk) RET(oz_slice_on(c, A, ld, T0, n_pad - T0, T0, nbT, oz::SCALE_CHOL, oz::MASK_LOWER, c->ozL, c->st));   // planes of panel op.col
          CU(cudaEventRecord(c->ev_pool[2 * op.col], c->st));            // panel op.col is complete on this rank
          if (!mine) mark(6);
          RET(inverse_step(op.col));
          break;
        }
        case DIST_UPDATE_SIDE: {                                           // look-ahead: panels pbeg .. pbeg+pcnt-1 -> my column
          cudaStream_t side = op.stream ? c->st2 : c->st3;
          CU(cudaStreamWaitEvent(side, c->ev_pool[2 * (op.pbeg + op.pcnt - 1)], 0));
          int klen = op.pcnt * NBO;
          if (op.pbeg * NBO + klen > n_pad) klen = n_pad - op.pbeg * NBO;
          // Optional cut of the long-k chunk into launches of <= kchunk (GPSS_DIST_KCHUNK).  Measured at 8 GPUs, n = 50k:
          // potrf 219 / 219 / 223 / 226 ms for kchunk = inf / 8192 / 4096 / 2048 (profiles/r01_dist_kchunk_sweep_8gpu.log),
          // i.e. the critical path is NOT waiting for CTA slots held by long-lived bulk CTAs; default: one launch.
          for (int k0 = 0; k0 < klen; k0 += kchunk)
            RET(update(T0, nbT, op.pbeg * NBO + k0, (klen - k0 < kchunk) ? (klen - k0) : kchunk, side));
          CU(cudaEventRecord(c->ev_pool[2 * op.col + 1], side));
          break;
        }
      }
    }
    if (pipe) CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * (nblk_o - 1)], 0));   // the last panel has arrived on the comm stream
    if (fused && ozk && !u2_int8) {                                              // every plane cut on st6 is complete before anything that follows
      CU(cudaEventRecord(c->ev_unpacked, c->st6));
      CU(cudaStreamWaitEvent(c->st, c->ev_unpacked, 0));
    }
    if (trace) {
      CU(cudaStreamSynchronize(c->st));
      const char* names[7] = {"wait for look-ahead updates", "U2 (panel j-1 -> my column)", "panel factorisation", "pack", "broadcast (as root)",
                              "broadcast (as receiver, incl. waiting for the owner)", "unpack"};
      if (fused) {
        names[1] = "U2 rows of the diagonal block (the rest runs beside the chain)";
        names[2] = "diagonal-block chain + inverse";
        names[3] = "wait for U2 + full-height solve into the broadcast buffer";
      }
      double sum[7] = {0, 0, 0, 0, 0, 0, 0};
      int cnt[7] = {0, 0, 0, 0, 0, 0, 0};
      for (size_t i = 1; i < marks.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
        sum[marks[i].first] += ms;
        cnt[marks[i].first]++;
      }
      for (auto& m : marks) cudaEventDestroy(m.second);
      fprintf(stderr, "[gpss dist trace] rank %d of %d, n_pad %d:", me, P, n_pad);
      for (int k = 0; k < 7; k++) fprintf(stderr, " %s: %.1f ms / %d;", names[k], sum[k], cnt[k]);
      fprintf(stderr, "\n");
    }
  } else
  for (int t = 0; t < nblk_o; t++) {
    const int T0 = t * NBO;
    const int nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    if (!la) {
      if (t >= 1) RET(update(T0, nbT, 0, T0, c->st));
      RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag));
      continue;
    }
    cudaEvent_t evP = c->ev_pool[2 * t], evU = c->ev_pool[2 * t + 1];
    if (t >= 2) CU(cudaStreamWaitEvent(c->st, evU, 0));              // U1(t) was issued on the side stream below
    if (t >= 1) RET(update(T0, nbT, T0 - NBO, NBO, c->st));          // U2(t)
    if (fused1) {                                                    // GPSS_PANEL=fused: the distributed path's panel on one GPU (tests, A/B timing)
      const long rows = n_pad - T0;
      RET(potrf_diag_chain(c, A, ld, T0, nbT, Winv, logdet_parts, dflag));
      if (rows > nbT) {
        GemmArgs g = gemm_args(A + (long)T0 * ld + T0 + nbT, ld, c->Wpan + (size_t)NBO * NBO, NBO, c->stage, rows - nbT, (int)rows - nbT, nbT, nbT);
        g.kend_col = 1;
        RET(gemm_ws_on(c, g, c->st));
        copy2d_kernel<<<592, 256, 0, c->st>>>(A + (long)T0 * ld + T0 + nbT, ld, c->stage, rows - nbT, rows - nbT, nbT);
        c->launches++;
      }
    } else
    RET(potrf_panel(c, A, ld, n_pad, T0, nbT, Winv, logdet_parts, dflag));
    if (ozk) RET(oz_slice_on(c, A, ld, T0, n_pad - T0, T0, nbT, oz::SCALE_CHOL, oz::MASK_LOWER, c->ozL, c->st));   // planes of panel t
    CU(cudaEventRecord(evP, c->st));
    // issue U1(t+1) = panels 0..t-1 applied to block column t+1; needs panel t-1 (complete: main stream order) --
    // here, right after panel t was ENQUEUED, the side stream must only wait for panel t-1.
    if (t + 1 < nblk_o && t >= 1) {
      const int T1 = T0 + NBO;
      const int nb1 = (n_pad - T1 < NBO) ? (n_pad - T1) : NBO;
      cudaStream_t side = (t & 1) ? c->st2 : c->st3;
      CU(cudaStreamWaitEvent(side, c->ev_pool[2 * (t - 1)], 0));
      RET(update(T1, nb1, 0, T0, side));
      CU(cudaEventRecord(c->ev_pool[2 * (t + 1) + 1], side));
    }
  }
  if (P > 1) {   // a failed pivot anywhere must be seen everywhere
    NC(g_nccl.AllReduce(dflag, dflag, 1, ncclInt, ncclMax, c->comm, c->st));
  }
  if (inter) c->trtri_inflight = true;                       // st8 / st9 still hold work: trtri_upper (or the next factorisation) joins them
  return GPSS_OK;
}

// ---------------------------------------------------------------------------------------------------
// PARTITIONED storage (n_pad^2 too large to replicate, e.g. n = 200 000: 320 GB): every rank keeps only the block columns
// it owns (j % P == rank), packed side by side (40 GB per rank at n = 200 000, P = 8).  RIGHT-looking factorisation: the
// owner factors block column t and broadcasts it; the broadcast buffer itself is the GEMM operand with which every rank
// updates its own remaining block columns (one k = 512 DMMA launch per panel over all of them, lower-triangle tiles only
// through the cyclic column map of gemm_nt_ws_kernel).  Look-ahead of one panel: the owner of t+1 updates that single
// column on the high-priority stream, factors and broadcasts it while the bulk update with panel t is still running.
// ---------------------------------------------------------------------------------------------------
static int potrf_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad;
  const long ld = n_pad;
  double* A = c->Lm;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  while ((int)c->ev_pool.size() < 2 * nblk_o + 2) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_pool.push_back(e);
  }
  const size_t per = (size_t)n_pad * NBO + (size_t)(NBO / NB) * NB * NB + NBO / NB;
  RET(ensure_stage(c, 3 * per));                                         // panels t, t-1, t-2 stay live (see the look-ahead below)
  // update of my local block columns [q0, q0 + cnt) with panel t, which lies in its broadcast buffer (rows T0.., ld = rows)
  auto update = [&](int t, int q0, int cnt, cudaStream_t stream) -> int {
    if (cnt <= 0) return GPSS_OK;
    const double* pan = c->stage + (size_t)(t % 3) * per;
    const int T0 = t * NBO, nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    const long rows = n_pad - T0;
    const int Rb = (q0 * P + me) * NBO;                                  // first global row (= first global column) touched
    long ncols = (long)cnt * NBO;
    if ((long)q0 * NBO + ncols > c->lcols) ncols = c->lcols - (long)q0 * NBO;   // ragged last block column
    GemmArgs g = gemm_args(pan + (Rb - T0), rows, pan, rows, A + (long)q0 * NBO * ld + Rb, ld, n_pad - Rb, (int)ncols, nbT);
    g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = Rb;
    g.cyc_P = P; g.cyc_me = me; g.cyc_w = NBO; g.cyc_lcol0 = q0 * NBO; g.cyc_boff = T0;
    return gemm_ws_on(c, g, stream);
  };
  // Look-ahead: the bulk update with panel s (side stream) covers my block columns j >= s + 3 only; column j receives
  // panels j-2 and j-1 on the MAIN stream when panel j-1 arrives.  The critical path (two k = 512 updates of one column,
  // the panel factorisation, the broadcast) therefore waits for the bulk update that finished a whole panel period
  // earlier (s = j - 3), never for the one in flight.
  for (int t = 0; t < nblk_o; t++) {
    const int T0 = t * NBO, nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    const bool mine = (t % P) == me;
    double* buf = c->stage + (size_t)(t % 3) * per;
    const long rows = n_pad - T0;
    const size_t n_panel = (size_t)rows * nbT, n_w = (size_t)(nbT / NB) * NB * NB, n_l = nbT / NB;
    double* Wt = c->Winv + (size_t)(T0 / NB) * NB * NB;
    // bulk update t-3 read this buffer and was the last side-stream launch to write block column t
    if (t >= 3) CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * (t - 3) + 1], 0));
    if (mine) {
      const int q = t / P;
      if (t >= 2) RET(update(t - 2, q, 1, c->st));
      if (t >= 1) RET(update(t - 1, q, 1, c->st));
      double* Acol = A + (long)q * NBO * ld;                              // my packed copy of global block column t
      RET(potrf_panel(c, Acol - (long)T0 * ld, ld, n_pad, T0, nbT, c->Winv, c->logdet_parts, c->dflag));   // indexes by global column
      pack_kernel<<<592, 256, 0, c->st>>>(buf, Acol + T0, ld, rows, nbT);
      CU(cudaMemcpyAsync(buf + n_panel, Wt, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
      CU(cudaMemcpyAsync(buf + n_panel + n_w, c->logdet_parts + T0 / NB, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
      c->launches++;
    }
    NC(g_nccl.Broadcast(buf, buf, n_panel + n_w + n_l, ncclDouble, t % P, c->comm, c->st));
    if (!mine) {
      CU(cudaMemcpyAsync(Wt, buf + n_panel, n_w * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
      CU(cudaMemcpyAsync(c->logdet_parts + T0 / NB, buf + n_panel + n_w, n_l * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
    }
    CU(cudaEventRecord(c->ev_pool[2 * t], c->st));
    int q0 = 0;                                                          // my first block column j >= t + 3
    while (q0 < c->nq && q0 * P + me < t + 3) q0++;
    CU(cudaStreamWaitEvent(c->st2, c->ev_pool[2 * t], 0));
    RET(update(t, q0, c->nq - q0, c->st2));
    CU(cudaEventRecord(c->ev_pool[2 * t + 1], c->st2));
  }
  CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * (nblk_o - 1) + 1], 0));
  NC(g_nccl.AllReduce(c->dflag, c->dflag, 1, ncclInt, ncclMax, c->comm, c->st));
  return GPSS_OK;
}

// alpha = L^-T L^-1 rhs with the partitioned factor.  Forward: the owner of block column t runs its four 128-steps and
// broadcasts the updated tail of the right-hand side and the finished piece of z.  Backward: every rank keeps the
// right-hand side current at the columns it owns and updates them with each new x_k; the owner of tile k-1 produces
// x_{k-1}, broadcast 128 doubles at a time.  rhs in c->rvec (destroyed), result in c->alpha (replicated).
static int potrs_vec_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad, nblk = c->nblk;
  const long ld = n_pad;
  const int w = NBO / NB;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  if (me == 0) { trsv_fwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv, c->rvec, c->zvec); c->launches++; }
  for (int t = 0; t < nblk_o; t++) {
    const int T0 = t * NBO, nbT = (n_pad - T0 < NBO) ? (n_pad - T0) : NBO;
    if ((t % P) == me) {
      const double* Lg = c->Lm + (long)(t / P) * NBO * ld - (long)T0 * ld;     // indexed by global column inside my block column
      for (int k = T0 / NB; k < (T0 + nbT) / NB && k + 1 < nblk; k++) {
        trsv_fwd_step_kernel<<<nblk - 1 - k, TRSV_THREADS, 0, c->st>>>(Lg, ld, c->Winv, c->rvec, c->zvec, k * NB);
        c->launches++;
      }
    }
    const int zend = (T0 + nbT + NB <= n_pad) ? T0 + nbT + NB : n_pad;          // z of this block column and of the next tile
    NC(g_nccl.Broadcast(c->zvec + T0, c->zvec + T0, (size_t)(zend - T0), ncclDouble, t % P, c->comm, c->st));
    if (T0 + nbT < n_pad)
      NC(g_nccl.Broadcast(c->rvec + T0 + nbT, c->rvec + T0 + nbT, (size_t)(n_pad - T0 - nbT), ncclDouble, t % P, c->comm, c->st));
  }
  CU(cudaGetLastError());
  const int own_last = ((nblk - 1) / w) % P;
  if (me == own_last) {
    trsv_bwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv + (long)(nblk - 1) * NB * NB, c->zvec, c->alpha, (nblk - 1) * NB);
    c->launches++;
  }
  NC(g_nccl.Broadcast(c->alpha + (long)(nblk - 1) * NB, c->alpha + (long)(nblk - 1) * NB, NB, ncclDouble, own_last, c->comm, c->st));
  const int ltiles = (int)(c->lcols / NB);
  for (int k = nblk - 1; k >= 1; k--) {
    trsv_bwd_step_part_kernel<<<ltiles, TRSV_THREADS, 0, c->st>>>(c->Lm, ld, c->Winv, c->zvec, c->alpha, k * NB, P, me, w);
    c->launches++;
    const int owner = ((k - 1) / w) % P;
    NC(g_nccl.Broadcast(c->alpha + (long)(k - 1) * NB, c->alpha + (long)(k - 1) * NB, NB, ncclDouble, owner, c->comm, c->st));
  }
  CU(cudaGetLastError());
  return GPSS_OK;
}

