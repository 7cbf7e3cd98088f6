// gpss_inverse.cuh -- part of the single translation unit gpss_capi.cu: U = L^-T, B^-1 = U U^T and the vector solves (single GPU
// and replicated layout), and the partitioned-storage inverse + gradient.
#pragma once
// ---------------------------------------------------------------------------------------------------
// PARTITIONED storage, gradient: B^-1 = U U^T with U = L^-T held as block ROWS owned cyclically (rank r keeps rows
// I = q P + r, packed: local row block q), 40 GB per rank at n = 200 000 like L itself.
//   inverse (left-looking over block columns J of U, as trtri_upper):
//       every rank contributes its blocks of the ROW strip L[J, 0:J] (ncclAllGather, then laid out in global column order);
//       the owner of J inverts the diagonal block and broadcasts W_JJ = inv(L_JJ);
//       T = U_loc[rows < J, 0:J0] L[J, 0:J0]^T  (k from each row's own start: cyclic row map of the GEMM kernel, split-k for
//       short slices),  U_loc[rows < J, J] = -T W_JJ^T.
//   B^-1 and the gradient, one block column J at a time: the owner broadcasts the row strip U[J, J0:], every rank forms
//       Q[I >= J, J] = U_loc[I, J0:] U[J, J0:]^T for its rows and feeds the 512-wide strip straight into the fused gradient
//       reductions -- B^-1 is never stored.
// Everything on the main stream; results (13 sums) all-reduced at the end.
// ---------------------------------------------------------------------------------------------------
static int trtri_diag_block(gpss_ctx* c, double* U, long ldu, const double* L, long ldl, int J0, int nbj);
static int pick_ksplit(int tiles, int klen, long part_doubles, size_t cap_doubles);
static int ensure_lazy(double** p, size_t count);

static int part_buffers(gpss_ctx* c)
{
  const size_t ldu = (size_t)(c->nq > 0 ? c->nq : 1) * NBO;
  RET(ensure_lazy(&c->Um, ldu * c->n_pad));
  RET(ensure_lazy(&c->Tpanel, ldu * NBO));                       // T, later the Q strip
  RET(ensure_lazy(&c->Wjj, (size_t)2 * NBO * NBO));              // W_JJ and the owner's U_JJ scratch
  const int nblk_o = (c->n_pad + NBO - 1) / NBO;
  const size_t cmax = (size_t)(nblk_o + c->world - 1) / c->world;
  RET(ensure_lazy(&c->pgather, (size_t)(c->world + 1) * cmax * NBO * NBO));   // [my piece | P gathered pieces]
  if (!c->Tsplit) {
    const size_t cap = (size_t)24576 * NBO;
    CU(cudaMalloc(&c->Tsplit, cap * sizeof(double)));
    c->Tsplit_cap = cap;
  }
  const long nblocks = (long)(ldu / NB) * c->nblk;               // every (local row tile, global column tile)
  if (c->partial_blocks < nblocks || !c->partial) {
    if (c->partial) cudaFree(c->partial);
    c->partial = nullptr;
    CU(cudaMalloc(&c->partial, sizeof(double) * (nblocks > 0 ? nblocks : 1) * NGRAD));
    c->partial_blocks = nblocks;
  }
  return GPSS_OK;
}

static int trtri_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad;
  const long ld = n_pad;
  const long ldu = (long)(c->nq > 0 ? c->nq : 1) * NBO;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  const int cmax = (nblk_o + P - 1) / P;
  const size_t blk = (size_t)NBO * NBO;
  double* piece = c->pgather;                                     // my blocks of the row strip
  double* gathered = c->pgather + (size_t)cmax * blk;             // P segments of cmax blocks
  double* Lrow = c->stage;                                        // the strip in global column order (the panel buffers are free now)
  double* Wjj = c->Wjj;
  double* Ujj = c->Wjj + blk;
  CU(cudaMemsetAsync(c->Um, 0, sizeof(double) * (size_t)ldu * n_pad, c->st));
  for (int J = 0; J < nblk_o; J++) {
    const int J0 = J * NBO, nbj = (n_pad - J0 < NBO) ? (n_pad - J0) : NBO;
    const int owner = J % P;
    if (owner == me) {
      // U_JJ from my packed copy of block column J (global addressing through shifted base pointers), W_JJ = U_JJ^T
      const double* Lg = c->Lm + (long)(J / P) * NBO * ld - (long)J0 * ld;
      double* Ug = Ujj - ((long)J0 * NBO + J0);
      CU(cudaMemsetAsync(Ujj, 0, sizeof(double) * blk, c->st));
      RET(trtri_diag_block(c, Ug, NBO, Lg, ld, J0, nbj));
      transpose_kernel<<<dim3(nbj / 32, nbj / 32), 256, 0, c->st>>>(Wjj, NBO, Ujj, NBO, 1);
      c->launches++;
      CU(cudaGetLastError());
    }
    NC(g_nccl.Broadcast(Wjj, Wjj, blk, ncclDouble, owner, c->comm, c->st));
    int cnt = 0;                                                  // my row blocks above J = my column blocks left of J
    while (cnt < c->nq && cnt * P + me < J) cnt++;
    if (J > 0) {
      const int jc = (J + P - 1) / P;                             // blocks per segment needed for this J (<= cmax)
      if (cnt > 0) {
        pack_rowstrip_kernel<<<592, 256, 0, c->st>>>(piece, c->Lm, ld, J0, nbj, NBO, cnt);
        c->launches++;
      }
      NC(g_nccl.AllGather(piece, gathered, (size_t)jc * blk, ncclDouble, c->comm, c->st));
      order_rowstrip_kernel<<<592, 256, 0, c->st>>>(Lrow, gathered, nbj, NBO, J, P, (long)jc * (long)blk);
      c->launches++;
      CU(cudaGetLastError());
    }
    if (cnt > 0 && J > 0) {
      const int rows = cnt * NBO;
      // T = U_loc[0:rows, 0:J0] * Lrow^T, k from each row's own global start
      GemmArgs g = gemm_args(c->Um, ldu, Lrow, nbj, c->Tpanel, ldu, rows, nbj, J0);
      g.kbeg_row = 1; g.rcyc_P = P; g.rcyc_me = me; g.rcyc_w = NBO; g.rcyc_l0 = 0; g.rcyc_koff = 0;
      const int S = pick_ksplit(rows / GemmTileWideWS::BM * (nbj / GemmTileWideWS::BN), J0, (long)rows * nbj, c->Tsplit_cap);
      if (S > 1) {
        g.C = c->Tsplit; g.ldc = rows; g.ksplit = S; g.csplit = (long)rows * nbj;
        RET(gemm(c, g));
        split_sum_kernel<<<296, 256, 0, c->st>>>(c->Tpanel, ldu, c->Tsplit, rows, nbj, S);
        c->launches++;
        CU(cudaGetLastError());
      } else {
        RET(gemm(c, g));
      }
      GemmArgs g2 = gemm_args(c->Tpanel, ldu, Wjj, NBO, c->Um + (long)J0 * ldu, ldu, rows, nbj, nbj);
      g2.negate_out = 1; g2.kend_col = 1;
      RET(gemm(c, g2));
    }
    if (owner == me) {                                            // my diagonal block
      copy2d_kernel<<<64, 256, 0, c->st>>>(c->Um + (long)J0 * ldu + (long)(J / P) * NBO, ldu, Ujj, NBO, nbj, nbj);
      c->launches++;
      CU(cudaGetLastError());
    }
  }
  return GPSS_OK;
}

static int gradient_partitioned(gpss_ctx* c)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad;
  const long ldu = (long)(c->nq > 0 ? c->nq : 1) * NBO;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  const int ltiles = (int)(ldu / NB);                             // my local row tiles (128 high)
  const int w = NBO / NB;
  double* strip = c->stage;                                       // U[J, J0:] as nbj x (n_pad - J0), contiguous
  double* Qs = c->Tpanel;                                         // Q[my rows >= J, J], ld = ldu
  const long nblocks = (long)ltiles * c->nblk;
  CU(cudaMemsetAsync(c->partial, 0, sizeof(double) * nblocks * NGRAD, c->st));
  for (int J = 0; J < nblk_o; J++) {
    const int J0 = J * NBO, nbj = (n_pad - J0 < NBO) ? (n_pad - J0) : NBO;
    const int owner = J % P;
    const long cols = n_pad - J0;
    if (owner == me) {
      pack_kernel<<<592, 256, 0, c->st>>>(strip, c->Um + (long)J0 * ldu + (long)(J / P) * NBO, ldu, nbj, cols);
      c->launches++;
    }
    NC(g_nccl.Broadcast(strip, strip, (size_t)nbj * cols, ncclDouble, owner, c->comm, c->st));
    int q0 = 0;                                                   // my first row block I >= J
    while (q0 < c->nq && q0 * P + me < J) q0++;
    const int rows = (c->nq - q0) * NBO - ((q0 < c->nq && (c->nq - 1) * P + me == nblk_o - 1) ? (NBO - (n_pad - (nblk_o - 1) * NBO)) : 0);
    if (rows <= 0) continue;
    // Q strip = U_loc[q0 rows.., J0:] * strip^T, k from each row's own start (relative to J0)
    GemmArgs g = gemm_args(c->Um + (long)J0 * ldu + (long)q0 * NBO, ldu, strip, nbj, Qs, ldu, rows, nbj, (int)cols);
    g.kbeg_row = 1; g.rcyc_P = P; g.rcyc_me = me; g.rcyc_w = NBO; g.rcyc_l0 = q0 * NBO; g.rcyc_koff = J0;
    RET(gemm(c, g));
    // fused gradient reductions over the strip: local row tiles q0*w .., global column tiles J*w ..
    const int ntm = rows / NB, ntn = nbj / NB;
    grad_pass_kernel<<<dim3(ntm, ntn), 256, 0, c->st>>>(Qs, ldu, c->zs, n_pad, c->xs, n_pad, c->alpha, c->n, c->dP,
                                                       c->partial + (long)(J * w) * ltiles * NGRAD, q0 * w, J * w, P, me, w);
    c->launches++;
    CU(cudaGetLastError());
  }
  sum_partials_kernel<NGRAD><<<1, 256, 0, c->st>>>(c->partial, nblocks, c->red + 8);
  c->launches++;
  CU(cudaGetLastError());
  NC(g_nccl.AllReduce(c->red + 8, c->red + 8, NGRAD, ncclDouble, ncclSum, c->comm, c->st));
  return GPSS_OK;
}

// PARTITIONED storage, predictive variance:  var_j = kD - || L^-1 b_j ||^2 = kD - || U^T b_j ||^2  with b_j = Sw k(X, x*_j) (GP_Utils.cpp:985-998).
// U = L^-T lives as cyclic block ROWS (trtri_partitioned); v = U^T b is needed by COLUMNS of U: rank r accumulates the block columns K it
// owns, v[K] = sum_{J <= K} U[J, K]^T b[J].  For every block row J the owner broadcasts the strip U[J, J0:] (transposed on the way, so
// that it is the NT product's B operand), and every rank adds its columns' share with ONE k = 512 DMMA launch over all of them (the cyclic
// column map of gemm_nt_ws_kernel, as in the partitioned Cholesky).  Then a sum of squares over the rank's columns and one all-reduce of
// m doubles.  Bm: the scaled cross-covariance batch, m_pad x n_pad, test index contiguous; result: c->dvar[0..mb) = || U^T b_j ||^2.
// The whole of U crosses NVLink once per batch of PRED_BATCH test points (4 n^2 bytes): a first version, sized for config 5's use
// (a trained n = 200 000 model queried at ~10^4 points), not for block models.
static int variance_partitioned(gpss_ctx* c, int mb, int m_pad)
{
  const int P = c->world, me = c->rank, n_pad = c->n_pad;
  const long ldu = (long)(c->nq > 0 ? c->nq : 1) * NBO;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  double* strip = c->stage;                                       // U[J, J0:]^T as (n_pad - J0) x nbj, contiguous
  double* Vt = c->Vm;                                             // -(U^T b)^T for my block columns: m_pad x lcols
  RET(ensure_stage(c, (size_t)n_pad * NBO));
  CU(cudaMemsetAsync(Vt, 0, sizeof(double) * (size_t)m_pad * c->lcols, c->st));
  for (int J = 0; J < nblk_o; J++) {
    const int J0 = J * NBO, nbj = (n_pad - J0 < NBO) ? (n_pad - J0) : NBO;
    const int owner = J % P;
    const long cols = n_pad - J0;
    if (owner == me) {
      transpose_kernel<<<dim3(nbj / 32, (unsigned)(cols / 32)), 256, 0, c->st>>>(strip, cols, c->Um + (long)J0 * ldu + (long)(J / P) * NBO, ldu, 0);
      c->launches++;
      CU(cudaGetLastError());
    }
    NC(g_nccl.Broadcast(strip, strip, (size_t)nbj * cols, ncclDouble, owner, c->comm, c->st));
    int q0 = 0;                                                   // my first block column K >= J
    while (q0 < c->nq && q0 * P + me < J) q0++;
    const long ncols = c->lcols - (long)q0 * NBO;
    if (ncols <= 0) continue;
    GemmArgs g = gemm_args(c->Bm + (long)J0 * m_pad, m_pad, strip, cols, Vt + (long)q0 * NBO * m_pad, m_pad, m_pad, (int)ncols, nbj);
    g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1;               // Vt <- Vt - b[J]^T U[J, my columns]: the sign does not matter for the norm
    g.cyc_P = P; g.cyc_me = me; g.cyc_w = NBO; g.cyc_lcol0 = q0 * NBO; g.cyc_boff = J0;
    RET(gemm(c, g));
  }
  rowsumsq_kernel<<<(mb + 255) / 256, 256, 0, c->st>>>(Vt, m_pad, mb, c->lcols, c->dvar);
  c->launches++;
  CU(cudaGetLastError());
  NC(g_nccl.AllReduce(c->dvar, c->dvar, (size_t)mb, ncclDouble, ncclSum, c->comm, c->st));
  return GPSS_OK;
}

static int create_streams(gpss_ctx* c)
{
  int lo = 0, hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority (largest number), hi = greatest
  CU(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, hi));
  // the bulk updates of the Cholesky sit one level above the lowest priority when the device has one to spare, so that the inverse the
  // distributed factorisation issues beside them (st9, lowest) only takes the CTA slots they leave
  const int bulk = (hi <= lo - 2) ? lo - 1 : lo;
  CU(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, bulk));
  CU(cudaStreamCreateWithPriority(&c->st3, cudaStreamNonBlocking, bulk));
  CU(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
  return GPSS_OK;
}

static void destroy_streams(gpss_ctx* c)
{
  if (c->ev_main) cudaEventDestroy(c->ev_main);
  if (c->ev_side) cudaEventDestroy(c->ev_side);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  c->ev_pool.clear();
  for (auto e : c->ev_pipe) cudaEventDestroy(e);
  c->ev_pipe.clear();
  if (c->st4) cudaStreamDestroy(c->st4);
  if (c->st5) cudaStreamDestroy(c->st5);
  if (c->st6) cudaStreamDestroy(c->st6);
  if (c->ev_u2) cudaEventDestroy(c->ev_u2);
  if (c->ev_unpacked) cudaEventDestroy(c->ev_unpacked);
  if (c->st7) cudaStreamDestroy(c->st7);
  if (c->st8) cudaStreamDestroy(c->st8);
  if (c->st9) cudaStreamDestroy(c->st9);
  c->st8 = c->st9 = nullptr;
  if (c->ev_factored) cudaEventDestroy(c->ev_factored);
  if (c->ev_solved) cudaEventDestroy(c->ev_solved);
  c->st4 = c->st5 = c->st6 = c->st7 = nullptr;
  c->ev_u2 = c->ev_unpacked = c->ev_factored = c->ev_solved = nullptr;
  if (c->st2) cudaStreamDestroy(c->st2);
  if (c->st3) cudaStreamDestroy(c->st3);
  if (c->st) cudaStreamDestroy(c->st);
  c->ev_main = c->ev_side = nullptr;
  c->st = c->st2 = c->st3 = nullptr;
}

// Number of k-parts for a GEMM of `tiles` output tiles on 2 x 148 CTA slots: the smallest S whose CTA count fills
// whole waves best, subject to parts of >= 2048 in k and to the capacity of the partial-product buffer.
static int pick_ksplit(int tiles, int klen, long part_doubles, size_t cap_doubles)
{
  const int slots = 296;
  if (tiles <= 0 || tiles >= 4 * slots) return 1;
  int best = 1;
  double best_eff = 0.0;
  for (int S = 1; S <= 8; S++) {
    if (S > 1 && (klen / S < 2048 || (size_t)part_doubles * S > cap_doubles)) break;
    const int ctas = tiles * S;
    const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
    if (eff > best_eff + 0.03) { best_eff = eff; best = S; }
  }
  return best;
}

// U = L^-T (upper), block-column left-looking; only NT products (see gpss_gemm.cuh header):
//     U[J,J]   = inv(L[J,J])^T                       built from the stored 128x128 inverses      (main stream)
//     U[0:J,J] = -(U[0:J,0:J] L[J,0:J]^T) U[J,J]     one long-k GEMM + one k = NBO GEMM           (side stream)
// The diagonal blocks depend only on L, so the main stream produces them (latency-bound small launches) ahead of
// the side stream, which runs the bulk GEMMs back to back.
// The diagonal block U[J0:J0+nbj, J0:J0+nbj] = inv(L[J0.., J0..])^T in 128-steps from the stored 128 x 128 inverses (main stream).
// U and L are addressed by GLOBAL row / column (callers with packed storage pass suitably shifted base pointers).
static int trtri_diag_block(gpss_ctx* c, double* U, long ldu, const double* L, long ldl, int J0, int nbj, cudaStream_t sm)
{
  for (int i0 = J0; i0 < J0 + nbj; i0 += NB) {
    const double* Wi = c->Winv + (long)(i0 / NB) * NB * NB;
    put_transposed_block_kernel<<<dim3(NB / 32, NB / 32), 256, 0, sm>>>(U + (long)i0 * ldu + i0, ldu, Wi);
    c->launches++;
    CU(cudaGetLastError());
    const int mr = i0 - J0;
    if (mr > 0) {
      double* Uc = U + (long)i0 * ldu + J0;                 // U[J0:i0, i0:i0+128]
      GemmArgs g = gemm_args(U + (long)J0 * ldu + J0, ldu, L + (long)J0 * ldl + i0, ldl, Uc, ldu, mr, NB, mr);
      g.kbeg_row = 1;
      RET(gemm_ws_on(c, g, sm));
      // Uc <- -Uc Wi^T in place: columns 64..127 first (all 128 inputs), then 0..63 (inputs 0..63 only)
      GemmArgs g1 = gemm_args(Uc, ldu, Wi + 64, NB, Uc + 64 * ldu, ldu, mr, 64, NB);
      g1.negate_out = 1;
      RET(gemm_ws_on(c, g1, sm));
      GemmArgs g2 = gemm_args(Uc, ldu, Wi, NB, Uc, ldu, mr, 64, 64);
      g2.negate_out = 1;
      RET(gemm_ws_on(c, g2, sm));
    }
  }
  return GPSS_OK;
}
static int trtri_diag_block(gpss_ctx* c, double* U, long ldu, const double* L, long ldl, int J0, int nbj)
{
  return trtri_diag_block(c, U, ldu, L, ldl, J0, nbj, c->st);
}

// One block column of the inverse.  sm: stream of the small diagonal-block launches, ss: stream of the bulk products, evs: the events that
// order ss after sm (entry 2 t belongs to step t).  trtri_upper runs all steps on (st, st2) after the factorisation; the distributed int8
// Cholesky issues step t on (st8, st9) as soon as panel t is complete on the rank (potrf_blocked), so that the inverse fills the time the
// tensor pipe spends waiting for the next panel.
static int trtri_step(gpss_ctx* c, const TrtriRun& R, int t)
{
  const long ld = c->n_pad;
  const int n_pad = c->n_pad;
  double *L = c->Lm, *U = c->Um;
  const int R0 = c->urow0, R1 = c->urow1;
  const bool ozk = R.ozk;
  std::vector<cudaEvent_t>& ev = *R.evs;
  {
    const int J0 = t * NBO;
    const int nbj = (n_pad - J0 < NBO) ? (n_pad - J0) : NBO;
    double* Wjj = c->Wjj + (size_t)t * NBO * NBO;
    // (1) the diagonal NBO-block of U in 128-steps
    RET(trtri_diag_block(c, U, ld, L, ld, J0, nbj, R.sm));
    // rows of block column t this rank reads as digit planes later: its own rows above the block and its share of the diagonal block
    const int sr0 = R0, sr1 = (R1 < J0 + nbj) ? R1 : (J0 + nbj);
    if (ozk && R.ss2) {
      // TWO bulk streams: block column t runs on A = (t odd ? ss2 : ss).  Its long product is cut at k = J0 - NBO: the part below the cut
      // needs only block columns <= t - 2 of U, so it starts while the OTHER stream still finishes column t - 1 (its k = NBO product, the
      // DMMA product with W_JJ and the digit planes, ~0.3 ms during which the int8 pipe used to idle, and the tail wave of a narrow slice);
      // the k = NBO rest follows when column t - 1 has been cut.  ev[2 t + 1] = digit planes of block column t complete.
      cudaStream_t A = (t & 1) ? R.ss2 : R.ss;
      double* Tb = (t & 1) ? R.T2 : c->Tpanel;
      if (t == 0) {
        CU(cudaEventRecord(ev[0], R.sm));
        CU(cudaStreamWaitEvent(A, ev[0], 0));
        if (sr1 > sr0) RET(oz_slice_on(c, U, ld, sr0, sr1 - sr0, 0, nbj, oz::SCALE_UNIT, oz::MASK_UPPER, c->ozU, A));
        CU(cudaEventRecord(ev[1], A));
        return GPSS_OK;
      }
      transpose_kernel<<<dim3(nbj / 32, nbj / 32), 256, 0, R.sm>>>(Wjj, NBO, U + (long)J0 * ld + J0, ld, 1);
      c->launches++;
      CU(cudaGetLastError());
      CU(cudaEventRecord(ev[2 * t], R.sm));
      const int ra = R0, rb = (R1 < J0) ? R1 : J0;           // my rows above this block column
      if (rb <= ra) {
        CU(cudaStreamWaitEvent(A, ev[2 * t], 0));
        if (sr1 > sr0) RET(oz_slice_on(c, U, ld, sr0, sr1 - sr0, J0, nbj, oz::SCALE_UNIT, oz::MASK_UPPER, c->ozU, A));
        CU(cudaEventRecord(ev[2 * t + 1], A));
        return GPSS_OK;
      }
      oz::Args a;
      memset(&a, 0, sizeof a);
      a.C = Tb + ra; a.ldc = ld; a.m = rb - ra; a.n = nbj;
      a.a_row0 = ra; a.b_row0 = J0; a.kbeg_row = 1;
      a.sign = 1.0; a.a_kind = oz::SCALE_UNIT; a.b_kind = oz::SCALE_CHOL;
      const int kcut = J0 - NBO;                             // columns of U below the cut were cut into planes by steps <= t - 2
      const bool has_bulk = kcut > (ra & ~(oz::BK - 1));
      if (has_bulk) {
        if (t >= 2) CU(cudaStreamWaitEvent(A, ev[2 * (t - 2) + 1], 0));      // (same stream: in order anyway; it also covers t - 3, see above)
        a.k0 = 0; a.k1 = kcut; a.accumulate = 0;
        RET(oz_gemm_on(c, c->oz_tmU[0], c->oz_tmL[1], a, A, c->oz_s_grad));
      }
      CU(cudaStreamWaitEvent(A, ev[2 * (t - 1) + 1], 0));                    // block column t - 1 of U is in planes
      a.k0 = has_bulk ? kcut : 0; a.k1 = J0; a.accumulate = has_bulk ? 1 : 0;
      RET(oz_gemm_on(c, c->oz_tmU[0], c->oz_tmL[1], a, A, c->oz_s_grad));
      CU(cudaStreamWaitEvent(A, ev[2 * t], 0));                              // W_JJ
      GemmArgs g2 = gemm_args(Tb + ra, ld, Wjj, NBO, U + (long)J0 * ld + ra, ld, rb - ra, nbj, nbj);
      g2.negate_out = 1; g2.kend_col = 1;
      RET(gemm_ws_on(c, g2, A));
      RET(oz_slice_on(c, U, ld, sr0, sr1 - sr0, J0, nbj, oz::SCALE_UNIT, oz::MASK_UPPER, c->ozU, A));
      CU(cudaEventRecord(ev[2 * t + 1], A));
      return GPSS_OK;
    }
    if (t == 0) {
      if (ozk && sr1 > sr0) {                                // digit planes of block column 0 (only its diagonal block)
        CU(cudaEventRecord(ev[0], R.sm));
        CU(cudaStreamWaitEvent(R.ss, ev[0], 0));
        RET(oz_slice_on(c, U, ld, sr0, sr1 - sr0, 0, nbj, oz::SCALE_UNIT, oz::MASK_UPPER, c->ozU, R.ss));
      }
      return GPSS_OK;
    }
    // (2) W_JJ = U_JJ^T into this block's scratch
    transpose_kernel<<<dim3(nbj / 32, nbj / 32), 256, 0, R.sm>>>(Wjj, NBO, U + (long)J0 * ld + J0, ld, 1);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaEventRecord(ev[2 * t], R.sm));
    CU(cudaStreamWaitEvent(R.ss, ev[2 * t], 0));
    const int ra = R0, rb = (R1 < J0) ? R1 : J0;             // my rows above this block column
    if (rb <= ra) {                                          // no rows above this block column: only my share of the diagonal block
      if (ozk && sr1 > sr0) RET(oz_slice_on(c, U, ld, sr0, sr1 - sr0, J0, nbj, oz::SCALE_UNIT, oz::MASK_UPPER, c->ozU, R.ss));
      return GPSS_OK;
    }
    // (3) T[ra:rb] = U[ra:rb, 0:J0] * L[Jblk, 0:J0]^T      (k starts at each tile's own row: U is upper triangular)
    GemmArgs g = gemm_args(U + ra, ld, L + J0, ld, c->Tpanel + ra, ld, rb - ra, nbj, J0);
    g.kbeg_row = 1; g.krow_off = ra;
    // A row slice has few tiles per step (rank 0 of 8 at n = 50k: 17 x 8 = 136 for 296 CTA slots) and the steps are
    // sequential, so a distributed rank cuts the long k-range of every tile into S parts (one CTA each), sized to
    // fill whole waves; the parts are summed in a fixed order by split_sum_kernel.
    if (ozk) {
      oz::Args a;
      memset(&a, 0, sizeof a);
      a.C = c->Tpanel + ra; a.ldc = ld; a.m = rb - ra; a.n = nbj;
      a.a_row0 = ra; a.b_row0 = J0; a.k0 = 0; a.k1 = J0; a.kbeg_row = 1;
      a.accumulate = 0; a.sign = 1.0; a.a_kind = oz::SCALE_UNIT; a.b_kind = oz::SCALE_CHOL;
      // Issued inside the distributed Cholesky (R.ss == st9), a CTA of this product would hold its SM for up to ~2 ms (k up to n) while the
      // panel kernels of the critical path -- the diagonal-block kernel cannot share an SM with it -- wait for one to drain.  GPSS_INV_KCHUNK
      // cuts the k-range into launches of that length (later ones accumulate), so SMs come free every fraction of a millisecond.
      const int inv_kchunk = [] { const char* e = getenv("GPSS_INV_KCHUNK"); const int v = e ? atoi(e) : 0; return v >= 1024 ? (v / 64) * 64 : 0; }();
      if (inv_kchunk > 0 && R.ss == c->st9 && J0 > inv_kchunk) {
        for (int k0 = (ra / inv_kchunk) * inv_kchunk; k0 < J0; k0 += inv_kchunk) {     // tiles start at their own row: nothing before ra
          oz::Args ac = a;
          ac.k0 = k0; ac.k1 = (k0 + inv_kchunk < J0) ? k0 + inv_kchunk : J0;
          ac.accumulate = (k0 > (ra / inv_kchunk) * inv_kchunk) ? 1 : 0;
          RET(oz_gemm_on(c, c->oz_tmU[0], c->oz_tmL[1], ac, R.ss, c->oz_s_grad));
        }
      } else
      RET(oz_gemm_on(c, c->oz_tmU[0], c->oz_tmL[1], a, R.ss, c->oz_s_grad));
      GemmArgs g2 = gemm_args(c->Tpanel + ra, ld, Wjj, NBO, U + (long)J0 * ld + ra, ld, rb - ra, nbj, nbj);
      g2.negate_out = 1; g2.kend_col = 1;
      RET(gemm_ws_on(c, g2, R.ss));
      RET(oz_slice_on(c, U, ld, sr0, sr1 - sr0, J0, nbj, oz::SCALE_UNIT, oz::MASK_UPPER, c->ozU, R.ss));   // planes of block column t
      return GPSS_OK;
    }
    int S = 1;
    if (c->world > 1) S = pick_ksplit((rb - ra) / GemmTileWideWS::BM * (nbj / GemmTileWideWS::BN), J0 - ra, (long)(rb - ra) * nbj, c->Tsplit_cap);
    if (S > 1) {
      const int rows = rb - ra;
      g.C = c->Tsplit; g.ldc = rows; g.ksplit = S; g.csplit = (long)rows * nbj;
      RET(gemm_ws_on(c, g, R.ss));
      split_sum_kernel<<<296, 256, 0, R.ss>>>(c->Tpanel + ra, ld, c->Tsplit, rows, nbj, S);
      c->launches++;
      CU(cudaGetLastError());
    } else {
      RET(gemm_ws_on(c, g, R.ss));
    }
    // (4) U[ra:rb, Jblk] = -T * W_JJ^T
    GemmArgs g2 = gemm_args(c->Tpanel + ra, ld, Wjj, NBO, U + (long)J0 * ld + ra, ld, rb - ra, nbj, nbj);
    g2.negate_out = 1; g2.kend_col = 1;
    RET(gemm_ws_on(c, g2, R.ss));
  }
  return GPSS_OK;
}

static int trtri_upper(gpss_ctx* c)
{
  const int n_pad = c->n_pad;
  const int nblk_o = (n_pad + NBO - 1) / NBO;
  if (c->trtri_inflight) {
    // the distributed Cholesky issued every step on (st8, st9) while it ran (potrf_blocked): only the join is left
    c->trtri_inflight = false;
    CU(cudaEventRecord(c->ev_side, c->st9));
    CU(cudaStreamWaitEvent(c->st, c->ev_side, 0));
    CU(cudaEventRecord(c->ev_side, c->st8));
    CU(cudaStreamWaitEvent(c->st, c->ev_side, 0));
    return GPSS_OK;
  }
  while ((int)c->ev_pool.size() < 2 * nblk_o + 2) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_pool.push_back(e);
  }
  // Distributed: every row of U depends only on L and on the SAME row of earlier block columns, so a rank computes
  // the rows [urow0, urow1) of its balanced slice with no communication (the 128-step diagonal blocks, which every
  // rank needs as right factors, are cheap and computed redundantly).
  // int8 path (gpss_ozaki.cuh): the long-k product (3) reads digit planes of U (cut block column by block column on the
  // side stream, right after (4) has written the column) and of L (cut by the factorisation); (4) and the diagonal blocks stay DMMA
  const bool ozk = oz_active(c) && c->ozL && c->ozU && c->ozL_valid;
  c->ozU_valid = ozk;
  c->dmma_coresident = ozk && c->oz_s <= 7 && !(getenv("GPSS_DMMA_CORESIDENT") && atoi(getenv("GPSS_DMMA_CORESIDENT")) == 0);   // see gemm_ws_on
  // the side stream must not start before the factor is complete on the main stream
  CU(cudaEventRecord(c->ev_main, c->st));
  CU(cudaStreamWaitEvent(c->st2, c->ev_main, 0));
  // int8 path, OPT-IN (GPSS_INV_TWO_STREAMS=1): consecutive block columns alternate between the two look-ahead streams (trtri_step).  Measured
  // on one B200 (profiles/r02_inverse_two_streams.log): correct, but SLOWER -- trtri 421 vs 410 ms at n = 50 000, 32.0 vs 31.1 ms at 20 000: with
  // >= 1000 CTAs per launch there is no idle pipe to fill on one GPU, and the cut costs every tile a second epilogue.
  const bool two = ozk && c->st3 && c->Tpanel2 && getenv("GPSS_INV_TWO_STREAMS") && atoi(getenv("GPSS_INV_TWO_STREAMS")) != 0;
  if (two) CU(cudaStreamWaitEvent(c->st3, c->ev_main, 0));
  const TrtriRun R = {c->st, c->st2, &c->ev_pool, ozk, two ? c->st3 : nullptr, two ? c->Tpanel2 : nullptr};
  for (int t = 0; t < nblk_o; t++) RET(trtri_step(c, R, t));
  CU(cudaEventRecord(c->ev_side, c->st2));
  CU(cudaStreamWaitEvent(c->st, c->ev_side, 0));
  if (two) {
    CU(cudaEventRecord(c->ev_side, c->st3));
    CU(cudaStreamWaitEvent(c->st, c->ev_side, 0));
  }
  c->dmma_coresident = false;
  return GPSS_OK;
}

// Distributed: every rank has computed its rows of U; B^-1 = U U^T and W = U^T need all of them.  A rank's slice is sent WITHOUT the
// diagonal 512-blocks (every rank computed those itself) and without the zeros below them (uslice_copy_kernel): at 2 GPUs and n = 50 000
// that is 8.3 GB instead of 16.7 GB.  All ranks pack their own slice at once; the broadcasts follow each other on the main stream while a
// helper stream unpacks the previous slice into U (two receive buffers).  35 ms -> see profiles/ for the measured time.
static long uslice_count(int n_pad, int r0, int rows)
{
  long cnt = 0;
  for (int b = 0; b * NBO < n_pad; b++) {
    long len = (long)b * NBO - r0;
    len = len < 0 ? 0 : (len > rows ? rows : len);
    const int w = (n_pad - b * NBO < NBO) ? (n_pad - b * NBO) : NBO;
    cnt += len * w;
  }
  return cnt;
}

static int allgather_U(gpss_ctx* c)
{
  if (c->world == 1) return GPSS_OK;
  const long ld = c->n_pad;
  const int P = c->world, me = c->rank;
  std::vector<int> b;
  balanced_rows(c->n_pad, P, c->urow_kind, b);
  std::vector<long> cnt(P);
  long mx = 0;
  for (int k = 0; k < P; k++) { cnt[k] = uslice_count(c->n_pad, b[k], b[k + 1] - b[k]); mx = std::max(mx, cnt[k]); }
  if (mx == 0) return GPSS_OK;
  RET(ensure_stage(c, (size_t)3 * mx));                        // [my packed slice | receive buffer 0 | receive buffer 1]
  if (!c->st5) {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&c->st5, cudaStreamNonBlocking, hi));
  }
  if (!c->ev_u2) CU(cudaEventCreateWithFlags(&c->ev_u2, cudaEventDisableTiming));
  while ((int)c->ev_pool.size() < 2 * P + 2) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_pool.push_back(e);
  }
  double* mine = c->stage;
  if (cnt[me] > 0) {
    uslice_copy_kernel<<<1184, 256, 0, c->st>>>(mine, c->Um, ld, c->n_pad, b[me], b[me + 1] - b[me], NBO, 0);
    c->launches++;
  }
  CU(cudaEventRecord(c->ev_u2, c->st));                        // the helper stream starts after everything the main stream holds so far
  CU(cudaStreamWaitEvent(c->st5, c->ev_u2, 0));
  int nrecv = 0;
  for (int k = 0; k < P; k++) {
    if (cnt[k] == 0) continue;
    double* rbuf = c->stage + (size_t)(1 + (nrecv & 1)) * mx;
    if (k != me && nrecv >= 2) CU(cudaStreamWaitEvent(c->st, c->ev_pool[2 * ((nrecv - 2) & 1) + 1], 0));   // that buffer's previous slice is unpacked
    NC(g_nccl.Broadcast(k == me ? mine : rbuf, k == me ? mine : rbuf, (size_t)cnt[k], ncclDouble, k, c->comm, c->st));
    if (k != me) {
      CU(cudaEventRecord(c->ev_pool[2 * (nrecv & 1)], c->st));
      CU(cudaStreamWaitEvent(c->st5, c->ev_pool[2 * (nrecv & 1)], 0));
      uslice_copy_kernel<<<1184, 256, 0, c->st5>>>(rbuf, c->Um, ld, c->n_pad, b[k], b[k + 1] - b[k], NBO, 1);
      c->launches++;
      CU(cudaEventRecord(c->ev_pool[2 * (nrecv & 1) + 1], c->st5));
      nrecv++;
    }
  }
  CU(cudaEventRecord(c->ev_u2, c->st5));
  CU(cudaStreamWaitEvent(c->st, c->ev_u2, 0));
  CU(cudaGetLastError());
  return GPSS_OK;
}

// Q (lower) = U U^T = B^-1; a rank computes the rows [qrow0, qrow1) of its balanced slice (all rows when alone)
static int lauum_lower(gpss_ctx* c)
{
  const long ld = c->n_pad;
  const int q0 = c->qrow0, q1 = c->qrow1;
  if (q1 <= q0) return GPSS_OK;
  if (oz_active(c) && c->ozU && (c->ozU_valid || c->world > 1)) {   // int8 path: both operands are the digit planes of U
    oz::Args a;
    memset(&a, 0, sizeof a);
    // distributed: the slices of the other ranks arrived as FP64 (allgather_U); cut every row a tile of mine can meet
    if (c->world > 1) RET(oz_slice_on(c, c->Um, ld, 0, q1, 0, c->n_pad, oz::SCALE_UNIT, oz::MASK_UPPER, c->ozU, c->st));
    a.C = c->Qm + q0; a.ldc = ld; a.m = q1 - q0; a.n = q1;
    a.a_row0 = q0; a.b_row0 = 0; a.k0 = 0; a.k1 = c->n_pad; a.kbeg_row = 1;
    a.lower_only = 1; a.grow0 = q0; a.gcol0 = 0; a.accumulate = 0; a.sign = 1.0; a.a_kind = oz::SCALE_UNIT; a.b_kind = oz::SCALE_UNIT;
    return oz_gemm_on(c, c->oz_tmU[0], c->oz_tmU[1], a, c->st, c->oz_s_grad);
  }
  GemmArgs g = gemm_args(c->Um + q0, ld, c->Um, ld, c->Qm + q0, ld, q1 - q0, q1, c->n_pad);
  g.lower_only = 1; g.kbeg_row = 1; g.krow_off = q0; g.grow0 = q0; g.gcol0 = 0;
  return gemm(c, g);
}

// x = L^-T L^-1 rhs through the stored diagonal inverses; rhs in c->rvec (destroyed), result in c->alpha
static int potrs_vec(gpss_ctx* c)
{
  const long ld = c->n_pad;
  const int nblk = c->nblk;
  trsv_fwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv, c->rvec, c->zvec);
  c->launches++;
  for (int k = 0; k + 1 < nblk; k++) {
    trsv_fwd_step_kernel<<<nblk - 1 - k, TRSV_THREADS, 0, c->st>>>(c->Lm, ld, c->Winv, c->rvec, c->zvec, k * NB);
    c->launches++;
  }
  CU(cudaGetLastError());
  trsv_bwd_first_kernel<<<1, TRSV_THREADS, 0, c->st>>>(c->Winv + (long)(nblk - 1) * NB * NB, c->zvec, c->alpha, (nblk - 1) * NB);
  c->launches++;
  for (int k = nblk - 1; k >= 1; k--) {
    trsv_bwd_step_kernel<<<k, TRSV_THREADS, 0, c->st>>>(c->Lm, ld, c->Winv, c->zvec, c->alpha, k * NB);
    c->launches++;
  }
  CU(cudaGetLastError());
  return GPSS_OK;
}

// dst = src / sn2 (the factor comes from the device parameters, so the launch carries no theta-dependent argument and can sit in a graph)
