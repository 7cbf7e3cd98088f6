// FP64 tensor-core GEMM "NT" workhorse for sm_100a:  C(MxN) <- op( A(MxK) * B(NxK)^T ),  all column-major.
//
// Every O(n^3) step of the exact-GP path is phrased as this one contraction:
//   potrf trailing update  A22 -= L21 L21^T          (LAPACK dpotrf inside arma::chol, GP_Utils.cpp:881,903)
//   panel solve            L21  = A21 inv(L11)^T     (same)
//   triangular inverse     U    = L^-T, block-column left-looking   } replace the two n x n dtrtrs of
//   Q = B^-1 = U U^T       (lower triangle)                          } solve_chol, GP_Utils.cpp:1202-1205
//   predictive variance    V    = L^-1 (Sw o kX)     (GP_Utils.cpp:985-996)
//
// B200 has no FP64 kind in tcgen05/UMMA; the FP64 tensor path is the warp-level DMMA.8x8x4
// (mma.sync.m8n8k4.f64), measured at 37.1 TFLOP/s (profiles/r01_fp64_peak_microbench.txt) versus
// 33.7 TFLOP/s for DFMA.  Operand tiles are staged in shared memory by a 4-deep cp.async (LDGSTS)
// pipeline with a +4-double row pad, which makes every 64-bit fragment load bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpss {

enum : int {
  GEMM_INIT_ZERO   = 0,   // acc = 0
  GEMM_INIT_NEGC   = 1,   // acc = -C   (with NEGATE_OUT gives C <- C - A B^T without touching the main loop)
};

struct GemmArgs {
  const double* A; long lda;     // A(i,k) at A[i + k*lda]
  const double* B; long ldb;     // B(j,k) at B[j + k*ldb]
  double* C;       long ldc;     // C(i,j) at C[i + j*ldc]
  int M, N, K;                   // M % BM == 0, N % BN == 0, K % 16 == 0
  int init_mode;                 // GEMM_INIT_*
  int negate_out;                // store -acc
  int lower_only;                // skip tiles that lie entirely above the diagonal of the GLOBAL matrix
  int grow0, gcol0;              // global (row, col) of C(0,0), used by lower_only
  int kbeg_row;                  // k starts at the tile's first row   (A upper-triangular in (row,k))
  int krow_off;                  // global row of A(0,.) / C(0,.) when the operands are a row slice (added to the tile row for kbeg_row / kend_row)
  int kend_row;                  // k ends   at the tile's last row+1  (A lower-triangular in (row,k))
  int kend_col;                  // k ends   at the tile's last col+1  (B lower-triangular in (col,k))
  int rev_order;                 // schedule tiles with the longest k-range first
  int ksplit;                    // > 1: every tile's k-range is cut into ksplit equal parts, one CTA each (grid = mt*nt*ksplit);
  long csplit;                   //      part s stores its partial product at C + s*csplit (summed by the caller, fixed order)
  int cyc_P, cyc_me, cyc_w;      // cyc_P > 1: C holds only the block columns (cyc_w wide) that rank cyc_me of cyc_P owns, packed;
  int cyc_lcol0, cyc_boff;       //   local column cyc_lcol0 + col is GLOBAL column gc = ((l / w) P + me) w + l % w, which lower_only
                                 //   tests against the row and which selects the B rows: B(gc - cyc_boff, k)
  int rcyc_P, rcyc_me, rcyc_w;   // rcyc_P > 1: the ROWS of A / C are block rows (rcyc_w high) owned cyclically and packed; local row
  int rcyc_l0, rcyc_koff;        //   rcyc_l0 + row is GLOBAL row gr = ((l / w) P + me) w + l % w, and kbeg_row starts k at gr - rcyc_koff
  int mt, nt;                    // tile counts M/BM, N/BN (filled by the launcher)
};

// CTA rasterisation: the grid is 1-D and walks the output in super-columns of GEMM_RASTER_W tile columns,
// tile-column fastest.  A wave of 2 x 148 CTAs then covers ~37 tile rows x 8 tile columns, so each operand
// k-slab is fetched from DRAM once per wave and shared through the 126 MB L2 (the row-fastest order it
// replaces streamed the whole A panel once per tile column: ~2 TB/s of DRAM reads in the big updates).
constexpr int GEMM_RASTER_W = 8;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int BM_, int BN_, int WARPS_M_, int WARPS_N_, int MIN_CTAS_>
struct GemmTile {
  static constexpr int BM = BM_, BN = BN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int MIN_CTAS = MIN_CTAS_;
  static constexpr int BK = 16;
  static constexpr int STAGES = 4;
  static constexpr int LDAS = BM + 4;          // (k*LDAS + m) mod 16 distinct over a half-warp: conflict-free LDS.64
  static constexpr int LDBS = BN + 4;
  static constexpr int WTM = BM / WARPS_M;     // warp tile rows  (64)
  static constexpr int WTN = BN / WARPS_N;     // warp tile cols  (32)
  static constexpr int MI = WTM / 8;           // mma tiles along M per warp
  static constexpr int NI = WTN / 8;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * BK * (LDAS + LDBS) * sizeof(double);
};

template <class T>
__global__ void __launch_bounds__(T::THREADS, T::MIN_CTAS) gemm_nt_kernel(const GemmArgs g)
{
  constexpr int BM = T::BM, BN = T::BN, BK = T::BK, STAGES = T::STAGES;
  constexpr int LDAS = T::LDAS, LDBS = T::LDBS, MI = T::MI, NI = T::NI;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * BK * LDAS;

  int tm, tn;
  {
    const int per_group = GEMM_RASTER_W * g.mt;
    const int grp = blockIdx.x / per_group, rem = blockIdx.x - grp * per_group;
    const int w = min(GEMM_RASTER_W, g.nt - grp * GEMM_RASTER_W);
    tm = rem / w;
    tn = grp * GEMM_RASTER_W + (rem - tm * w);
  }
  if (g.rev_order) { tm = g.mt - 1 - tm; }
  const int row0 = tm * BM, col0 = tn * BN;
  if (g.lower_only && (g.gcol0 + col0) > (g.grow0 + row0 + BM - 1)) return;

  int kbeg = 0, kend = g.K;
  if (g.kbeg_row) kbeg = g.krow_off + row0;
  if (g.kend_row) kend = min(kend, g.krow_off + row0 + BM);
  if (g.kend_col) kend = min(kend, col0 + BN);
  kbeg = (kbeg / BK) * BK;
  const int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % T::WARPS_M) * T::WTM;
  const int wn0 = (warp / T::WARPS_M) * T::WTN;
  const int lr = lane >> 2, lc = lane & 3;

  const double* Ag = g.A + row0;
  const double* Bg = g.B + col0;

  auto load_stage = [&](int slot, int k0) {
    double* as = As + slot * BK * LDAS;
    double* bs = Bs + slot * BK * LDBS;
#pragma unroll
    for (int idx = tid; idx < BK * (BM / 2); idx += T::THREADS) {
      const int kk = idx / (BM / 2), c = idx % (BM / 2);
      cp_async16(as + kk * LDAS + 2 * c, Ag + (long)(k0 + kk) * g.lda + 2 * c);
    }
#pragma unroll
    for (int idx = tid; idx < BK * (BN / 2); idx += T::THREADS) {
      const int kk = idx / (BN / 2), c = idx % (BN / 2);
      cp_async16(bs + kk * LDBS + 2 * c, Bg + (long)(k0 + kk) * g.ldb + 2 * c);
    }
  };

  // start the operand pipeline first so the accumulator initialisation overlaps it
#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) load_stage(s, kbeg + s * BK);
    cp_async_commit();
  }

  double acc[MI][NI][2];
  double* Cg = g.C + (long)(col0 + wn0 + 2 * lc) * g.ldc + (row0 + wm0 + lr);
  if (g.init_mode == GEMM_INIT_NEGC) {
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) {
        acc[i][j][0] = -Cg[(long)(j * 8) * g.ldc + i * 8];
        acc[i][j][1] = -Cg[(long)(j * 8 + 1) * g.ldc + i * 8];
      }
  } else {
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  }

  for (int it = 0; it < nk; it++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = it + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, kbeg + nxt * BK);
      cp_async_commit();
    }
    const double* as = As + (it % STAGES) * BK * LDAS + wm0 + lr;
    const double* bs = Bs + (it % STAGES) * BK * LDBS + wn0 + lr;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      double a[MI], b[NI];
      const int kq = k4 * 4 + lc;
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = as[kq * LDAS + i * 8];
#pragma unroll
      for (int j = 0; j < NI; j++) b[j] = bs[kq * LDBS + j * 8];
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  const double sgn = g.negate_out ? -1.0 : 1.0;
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) {
      Cg[(long)(j * 8) * g.ldc + i * 8] = sgn * acc[i][j][0];
      Cg[(long)(j * 8 + 1) * g.ldc + i * 8] = sgn * acc[i][j][1];
    }
}

// ---------------------------------------------------------------------------------------------------
// Warp-specialised variant: one PRODUCER warp feeds the operand ring with bulk asynchronous copies
// (cp.async.bulk global->shared, 1 KB / 512 B rows, completion counted on an mbarrier), the consumer
// warps only wait on the stage's "full" mbarrier, load fragments and issue DMMA, then release the stage
// on its "empty" mbarrier.  Compared with gemm_nt_kernel this removes from the DMMA warps all the
// per-stage work the ncu source view charged ~10% of the issue time to: the 64-bit address arithmetic of
// 12 LDGSTS per thread, LDGDEPBAR/DEPBAR and the CTA-wide BAR.SYNC (profiles/r01_gemm_wide_source.txt).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int BM_, int BN_, int WARPS_M_, int WARPS_N_, int MIN_CTAS_, int STAGES_ = 4>
struct GemmTileWS {
  static constexpr int BM = BM_, BN = BN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
  static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;      // + the producer warp
  static constexpr int MIN_CTAS = MIN_CTAS_;
  static constexpr int BK = 16;
  static constexpr int STAGES = STAGES_;
  static constexpr int LDAS = BM + 4;
  static constexpr int LDBS = BN + 4;
  static constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
  static constexpr int MI = WTM / 8, NI = WTN / 8;
  static constexpr uint32_t STAGE_TX_BYTES = (uint32_t)BK * (BM + BN) * sizeof(double);
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * BK * (LDAS + LDBS) * sizeof(double) + 2 * STAGES * sizeof(uint64_t);
};

template <class T>
__global__ void __launch_bounds__(T::THREADS, T::MIN_CTAS) gemm_nt_ws_kernel(const GemmArgs g)
{
  constexpr int BM = T::BM, BN = T::BN, BK = T::BK, STAGES = T::STAGES;
  constexpr int LDAS = T::LDAS, LDBS = T::LDBS, MI = T::MI, NI = T::NI;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * BK * LDAS;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * BK * (LDAS + LDBS));
  uint64_t* empty_bar = full_bar + STAGES;

  int tm, tn, split = 0;
  {
    int bid = blockIdx.x;
    if (g.ksplit > 1) { const int tiles = g.mt * g.nt; split = bid / tiles; bid -= split * tiles; }
    const int per_group = GEMM_RASTER_W * g.mt;
    const int grp = bid / per_group, rem = bid - grp * per_group;
    const int w = min(GEMM_RASTER_W, g.nt - grp * GEMM_RASTER_W);
    tm = rem / w;
    tn = grp * GEMM_RASTER_W + (rem - tm * w);
  }
  if (g.rev_order) { tm = g.mt - 1 - tm; }
  const int row0 = tm * BM, col0 = tn * BN;
  int bcol0 = col0, gcol = g.gcol0 + col0;
  if (g.cyc_P > 1) {
    const int l = g.cyc_lcol0 + col0;
    gcol = ((l / g.cyc_w) * g.cyc_P + g.cyc_me) * g.cyc_w + l % g.cyc_w;
    bcol0 = gcol - g.cyc_boff;
  }
  if (g.lower_only && gcol > (g.grow0 + row0 + BM - 1)) return;

  int kbeg = 0, kend = g.K;
  if (g.kbeg_row) {
    if (g.rcyc_P > 1) {
      const int l = g.rcyc_l0 + row0;
      kbeg = max(0, ((l / g.rcyc_w) * g.rcyc_P + g.rcyc_me) * g.rcyc_w + l % g.rcyc_w - g.rcyc_koff);
    } else {
      kbeg = g.krow_off + row0;
    }
  }
  if (g.kend_row) kend = min(kend, g.krow_off + row0 + BM);
  if (g.kend_col) kend = min(kend, col0 + BN);
  kbeg = (kbeg / BK) * BK;
  int nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  if (g.ksplit > 1) {   // this CTA's share of the tile's k-steps (possibly none: it then stores the initial accumulator)
    const int chunk = (nk + g.ksplit - 1) / g.ksplit;
    const int first = min(nk, split * chunk);
    nk = min(nk - first, chunk);
    kbeg += first * BK;
  }

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, T::CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == T::CONSUMER_WARPS) {
    // ---------------- producer warp: lanes 0..15 copy the A rows, lanes 16..31 the B rows of each stage ----------------
    const int kk = lane & 15;
    const bool isA = lane < 16;
    const double* src = isA ? (g.A + row0 + (long)(kbeg + kk) * g.lda) : (g.B + bcol0 + (long)(kbeg + kk) * g.ldb);
    const long src_step = (long)BK * (isA ? g.lda : g.ldb);
    double* dst0 = isA ? (As + kk * LDAS) : (Bs + kk * LDBS);
    const int dst_step = BK * (isA ? LDAS : LDBS);
    const uint32_t bytes = (uint32_t)((isA ? BM : BN) * sizeof(double));
    for (int it = 0; it < nk; it++) {
      const int slot = it % STAGES;
      if (it >= STAGES) mbar_wait(empty_bar + slot, ((it / STAGES) - 1) & 1);
      if (lane == 0) mbar_arrive_expect_tx(full_bar + slot, T::STAGE_TX_BYTES);
      __syncwarp();
      bulk_g2s(dst0 + slot * dst_step, src, bytes, full_bar + slot);
      src += src_step;
    }
    return;
  }

  // ---------------- consumer warps ----------------
  const int wm0 = (warp % T::WARPS_M) * T::WTM;
  const int wn0 = (warp / T::WARPS_M) * T::WTN;
  const int lr = lane >> 2, lc = lane & 3;

  double acc[MI][NI][2];
  double* Cg = g.C + (long)split * g.csplit + (long)(col0 + wn0 + 2 * lc) * g.ldc + (row0 + wm0 + lr);
  if (g.init_mode == GEMM_INIT_NEGC) {
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) {
        acc[i][j][0] = -Cg[(long)(j * 8) * g.ldc + i * 8];
        acc[i][j][1] = -Cg[(long)(j * 8 + 1) * g.ldc + i * 8];
      }
  } else {
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  }

  for (int it = 0; it < nk; it++) {
    const int slot = it % STAGES;
    mbar_wait(full_bar + slot, (it / STAGES) & 1);
    if (it > 0) {
      // Release the PREVIOUS stage here, not right after its last LDS: SYNCS.ARRIVE is not held back by LDS that
      // are still in flight (ptxas also hoists it above the DMMAs), and with a second, out-of-phase CTA on the SM
      // the producer's refill was observed to overtake them (bench_micro/gemm_stress.cu, "concurrent" cases).
      // At this point every DMMA of the previous stage has been issued, hence all its LDS have delivered.
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar + (it - 1) % STAGES);
    }
    const double* as = As + slot * BK * LDAS + wm0 + lr;
    const double* bs = Bs + slot * BK * LDBS + wn0 + lr;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; k4++) {
      double a[MI], b[NI];
      const int kq = k4 * 4 + lc;
#pragma unroll
      for (int i = 0; i < MI; i++) a[i] = as[kq * LDAS + i * 8];
#pragma unroll
      for (int j = 0; j < NI; j++) b[j] = bs[kq * LDBS + j * 8];
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }

  const double sgn = g.negate_out ? -1.0 : 1.0;
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) {
      Cg[(long)(j * 8) * g.ldc + i * 8] = sgn * acc[i][j][0];
      Cg[(long)(j * 8 + 1) * g.ldc + i * 8] = sgn * acc[i][j][1];
    }
}

using GemmTileWideWS = GemmTileWS<128, 64, 2, 2, 2>;
// the same tile with a 2-stage ring: 51 KB of shared memory, so that a CTA fits on an SM NEXT TO a resident oz_gemm_kernel CTA (gemm_ws_on)
using GemmTileWideWS2 = GemmTileWS<128, 64, 2, 2, 2, 2>;

// 128x64 tile, 4 warps (64x32 each), 2 CTAs per SM: one CTA's prologue/epilogue hides behind the other's DMMA loop.
using GemmTileWide = GemmTile<128, 64, 2, 2, 2>;

}  // namespace gpss
