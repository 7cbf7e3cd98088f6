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
  int kend_row;                  // k ends   at the tile's last row+1  (A lower-triangular in (row,k))
  int kend_col;                  // k ends   at the tile's last col+1  (B lower-triangular in (col,k))
  int rev_order;                 // schedule tiles with the longest k-range first
};

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

  int tm = blockIdx.x, tn = blockIdx.y;
  if (g.rev_order) { tm = gridDim.x - 1 - tm; }
  const int row0 = tm * BM, col0 = tn * BN;
  if (g.lower_only && (g.gcol0 + col0) > (g.grow0 + row0 + BM - 1)) return;

  int kbeg = 0, kend = g.K;
  if (g.kbeg_row) kbeg = row0;
  if (g.kend_row) kend = min(kend, row0 + BM);
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

// 128x64 tile, 4 warps (64x32 each), 2 CTAs per SM: one CTA's prologue/epilogue hides behind the other's DMMA loop.
using GemmTileWide = GemmTile<128, 64, 2, 2, 2>;
// 128x128 tile, 8 warps, 1 CTA per SM: a single CTA owns the full 128-wide panel row, which makes the
// in-place panel solve  A21 <- A21 inv(L11)^T  hazard free.
using GemmTilePanel = GemmTile<128, 128, 2, 4, 1>;

template <class T>
inline cudaError_t gemm_nt_launch(const GemmArgs& g, cudaStream_t st)
{
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  dim3 grid(g.M / T::BM, g.N / T::BN);
  gemm_nt_kernel<T><<<grid, T::THREADS, T::SMEM_BYTES, st>>>(g);
  return cudaGetLastError();
}

}  // namespace gpss
