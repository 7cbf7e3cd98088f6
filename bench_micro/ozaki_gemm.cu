// EXPERIMENTAL (round-2 candidate, NOT on the product path, NOT yet run on a GPU): the FP64 contraction
//     C <- C - A B^T          (A: m x k, B: n x k, C: m x n, all column-major FP64 like the rest of the path)
// evaluated on the int8 tensor cores of sm_100a (tcgen05.mma kind::i8, int32 accumulators in TMEM) with the Ozaki
// splitting instead of DMMA.  scripts/ozaki_numerics.py is the CPU study that fixes the numerics: with ONE a-priori
// power-of-two scale per operand (|L_ij| <= sqrt(max B_ii), |L^-1_ij| <= 1 on this path) s = 7 slices of 7 bits
// (28 int8 products) keep nlml / alpha / g two orders inside the parity tolerances, s = 8 (36 products) reaches
// the FP64 noise floor.  At the B200's int8 rate that is a 2-3x higher ceiling than the 37 TFLOP/s DMMA pipe.
//
//   operand x -> t = x / 2^e, v = rint(t 2^(7s-1)) (one rounding), signed base-128 digits d_0 .. d_{s-1} in [-64, 64]:
//       t = sum_p d_p 2^-(7p+6)                                   (oz_slice_kernel; planes are K-major int8)
//   A B^T = 2^(eA+eB-12) sum_g 2^(-7g) G_g,   G_g = sum_{i+j=g} A_i B_j^T  (exact in int32 for k <= 65 536)
//   pairs with i + j >= s are dropped (below the rounding of v).
//
// Kernel (oz_gemm_kernel<S>): one CTA per 128 x 64 tile of C, 192 threads:
//   warp 4   : TMA producer -- per 64-byte k-chunk all S planes of the A tile (128 rows) and of the B tile (64 rows),
//              SWIZZLE_64B boxes, one mbarrier per stage;
//   warp 5   : allocates TMEM (S x 64 columns), issues S (S + 1) / 2 x 2 tcgen05.mma (M 128, N 64, K 32) per chunk,
//              group g accumulates in TMEM columns [64 g, 64 g + 64); tcgen05.commit releases the stage;
//   warps 0-3: epilogue -- tcgen05.ld of the S group accumulators, Horner in FP64, C read-modify-write (coalesced:
//              TMEM lane = row of C).
// Every plane of a chunk is loaded ONCE and used by up to S products: 12 S bytes of L2 traffic per 64 x S (S+1) / 2
// x 8192 MACs ~ 48 B/clk/SM at the full MMA rate, close to the measured ~42 B/clk/SM L2 cap (B300_MICROARCH.md).
//
//   ./ozaki_gemm exact            int32 exactness of the MMA / descriptor / TMEM plumbing against a CPU integer GEMM
//   ./ozaki_gemm check [S]        FP64 result against a CPU long-double reference (512 x 512 x 1024)
//   ./ozaki_gemm bench [M N K S]  throughput, beside the DMMA kernel of the product on the same shape
#include "../gp_ss_ak_b200/csrc/gpss_gemm.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

using namespace gpss;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

namespace oz {

constexpr int BM = 128, BN = 64, BK = 64;        // BK in bytes (= int8 elements) per stage and plane: one SWIZZLE_64B row
constexpr int UMMA_K = 32;                       // kind::i8: 32 bytes of k per instruction
constexpr int DIGIT_BITS = 7;
constexpr int RASTER_W = 8;

template <int S>
struct Cfg {
  static constexpr int PAIRS = S * (S + 1) / 2;
  static constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK;
  static constexpr int STAGE_BYTES = S * (A_BYTES + B_BYTES);
  static constexpr int STAGES_FIT = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 4 ? 4 : STAGES_FIT;
  static constexpr int TMEM_COLS = S * BN <= 64 ? 64 : S * BN <= 128 ? 128 : S * BN <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
  static_assert(S >= 1 && S * BN <= 512, "S group accumulators of BN columns must fit the 512 TMEM columns");
  static_assert(STAGES >= 2, "need at least two stages");
};

struct Args {
  double* C; long ldc;           // C(i,j) at C[i + j*ldc], updated in place: C += scale * sum_g 2^(-7g) G_g
  int m, n, k;                   // m % 128 == 0, n % 64 == 0, k % 64 == 0
  int a_rows, b_rows;            // rows per plane of the plane tensors (plane p, row r -> TMA row p * rows + r)
  int a_row0, b_row0, k0;        // offsets of this product inside the plane tensors
  double scale;                  // -2^(eA + eB - 12) for C - A B^T
  int32_t* dbg;                  // exact mode: raw int32 group accumulators, [S][m][n] row-major (else nullptr)
};

// ------------------------------------------------------------------ PTX helpers (sm_100a)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
               :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar)       // arrives on bar when all MMAs issued so far have completed
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// K-major operand tile in shared memory, rows of 64 bytes, SWIZZLE_64B (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp:
// start >> 4 in [0,14), LBO in [16,30) (ignored for swizzled K-major: 1), SBO = 8 rows x 64 B >> 4 = 32 in [32,46),
// version 1 in [46,48), layout SWIZZLE_64B = 4 in [61,64)).  The tile base is 1024-byte aligned.
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr)
{
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format S32 = 2 [4,6), a/b_format INT8 = 1 [7,10) / [10,13), K-major both, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

template <int S>
__global__ void __launch_bounds__(192, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Args g)
{
  using T = Cfg<S>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + T::STAGES * T::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + T::STAGES;
  uint64_t* acc_bar = empty_bar + T::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rasterisation as in gemm_nt_ws_kernel: super-columns of RASTER_W tile columns, tile column fastest, so that one wave of
  // 148 CTAs shares ~19 A row-tiles x 8 B row-tiles through L2 instead of streaming every A plane once per tile column
  const int mt = g.m / BM, nt = g.n / BN;
  const int grp = blockIdx.x / (RASTER_W * mt), within = blockIdx.x % (RASTER_W * mt);
  const int gcols = min(RASTER_W, nt - grp * RASTER_W);
  const int tile_n = grp * RASTER_W + within % gcols, tile_m = within / gcols;
  const int nk = g.k / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < T::STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)T::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kc = 0; kc < nk; kc++) {
        const int st = kc % T::STAGES;
        const uint32_t ph = (uint32_t)(kc / T::STAGES) & 1u;
        mbar_wait(empty_bar + st, ph ^ 1u);
        mbar_arrive_expect_tx(full_bar + st, (uint32_t)T::STAGE_BYTES);
        uint8_t* sa = smem + st * T::STAGE_BYTES;
        uint8_t* sb = sa + S * T::A_BYTES;
        const int kb = g.k0 + kc * BK;
#pragma unroll
        for (int p = 0; p < S; p++) {
          tma_load_2d(sa + p * T::A_BYTES, &tmA, full_bar + st, kb, p * g.a_rows + g.a_row0 + tile_m * BM);
          tma_load_2d(sb + p * T::B_BYTES, &tmB, full_bar + st, kb, p * g.b_rows + g.b_row0 + tile_n * BN);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      for (int kc = 0; kc < nk; kc++) {
        const int st = kc % T::STAGES;
        const uint32_t ph = (uint32_t)(kc / T::STAGES) & 1u;
        mbar_wait(full_bar + st, ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + st * T::STAGE_BYTES);
        const uint32_t sb = sa + S * T::A_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / UMMA_K; ks++) {
#pragma unroll
          for (int i = 0; i < S; i++) {
            const uint64_t ad = smem_desc_sw64(sa + i * T::A_BYTES + ks * UMMA_K);
#pragma unroll
            for (int j = 0; j < S - i; j++) {
              const uint64_t bd = smem_desc_sw64(sb + j * T::B_BYTES + ks * UMMA_K);
              // group i + j: the first product that reaches it (i == 0 of the first k-step) overwrites, the rest accumulate
              mma_i8(tmem_base + (uint32_t)((i + j) * BN), ad, bd, IDESC_I8, (kc > 0 || ks > 0 || i > 0) ? 1u : 0u);
            }
          }
        }
        tc_commit(empty_bar + st);         // stage reusable once these MMAs have read it
      }
      tc_commit(acc_bar);                  // accumulators complete
    }
  } else {
    // ------------------------------------------------ epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) = rows of the tile
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int row = tile_m * BM + warp * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const double w = 1.0 / (double)(1 << DIGIT_BITS);
    for (int c0 = 0; c0 < BN; c0 += 8) {
      uint32_t v[S][8];
#pragma unroll
      for (int gi = 0; gi < S; gi++) tmem_ld8(lane_base + (uint32_t)(gi * BN + c0), v[gi]);
      tmem_ld_wait();
      if (g.dbg) {
#pragma unroll
        for (int gi = 0; gi < S; gi++)
#pragma unroll
          for (int c = 0; c < 8; c++) g.dbg[((size_t)gi * g.m + row) * g.n + tile_n * BN + c0 + c] = (int32_t)v[gi][c];
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double acc = 0.0;
#pragma unroll
        for (int gi = S - 1; gi >= 0; gi--) acc = acc * w + (double)(int32_t)v[gi][c];     // sum_g 2^(-7g) G_g, smallest first
        double* cp = g.C + row + (size_t)(tile_n * BN + c0 + c) * g.ldc;
        *cp = *cp + g.scale * acc;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"((uint32_t)T::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ slicing: FP64 column-major -> S K-major int8 planes
// X(r, kk) at X[r + kk*ldx], rows x k.  planes[p][(row0 + r) * kpad + k0 + kk] = digit p.  32 x 32 tiles through shared
// memory so both sides are coalesced (harness quality; the product version would pack 4 digits per store).
template <int S>
__global__ void oz_slice_kernel(const double* X, long ldx, int rows, int k, double inv_scale,
                                int8_t* planes, long plane_rows, long kpad, int row0, int k0)
{
  __shared__ int8_t tile[S][32][33];
  const int r = blockIdx.x * 32 + threadIdx.x;
  for (int ky = threadIdx.y; ky < 32; ky += blockDim.y) {
    const int kq = blockIdx.y * 32 + ky;
    int d[S];
#pragma unroll
    for (int p = 0; p < S; p++) d[p] = 0;
    if (r < rows && kq < k) {
      const double lim = (double)(1ll << (DIGIT_BITS * S - 1));
      double sc = X[r + (size_t)kq * ldx] * inv_scale * lim;          // exact: powers of two
      sc = fmin(fmax(sc, -lim), lim);
      long long v = __double2ll_rn(sc);
#pragma unroll
      for (int p = S - 1; p >= 1; p--) {
        const int dg = (int)((v + 64) & 127) - 64;                    // [-64, 63], exact remainder
        v = (v - dg) >> DIGIT_BITS;
        d[p] = dg;
      }
      d[0] = (int)v;                                                   // |v| <= 64
    }
#pragma unroll
    for (int p = 0; p < S; p++) tile[p][threadIdx.x][ky] = (int8_t)d[p];
  }
  __syncthreads();
  for (int ry = threadIdx.y; ry < 32; ry += blockDim.y) {
    const int rr = blockIdx.x * 32 + ry, kq = blockIdx.y * 32 + threadIdx.x;
    if (rr < rows && kq < k) {
#pragma unroll
      for (int p = 0; p < S; p++) planes[((size_t)p * plane_rows + row0 + rr) * kpad + k0 + kq] = tile[p][ry][threadIdx.x];
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); exit(1); }
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// planes: [S * plane_rows][kpad] bytes, K contiguous; box = 64 bytes of k x box_rows rows, SWIZZLE_64B
static CUtensorMap make_plane_map(const int8_t* planes, long total_rows, long kpad, int box_rows)
{
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {(cuuint64_t)kpad};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)planes, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return tm;
}

template <int S>
static void launch(const int8_t* pa, const int8_t* pb, Args g, long kpad, cudaStream_t st = 0)
{
  using T = Cfg<S>;
  static bool attr = false;
  if (!attr) { CK(cudaFuncSetAttribute(oz_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES)); attr = true; }
  CUtensorMap ta = make_plane_map(pa, (long)S * g.a_rows, kpad, BM);
  CUtensorMap tb = make_plane_map(pb, (long)S * g.b_rows, kpad, BN);
  oz_gemm_kernel<S><<<(g.m / BM) * (g.n / BN), 192, T::SMEM_BYTES, st>>>(ta, tb, g);
}

template <int S>
static void slice(const double* X, long ldx, int rows, int k, double inv_scale, int8_t* planes, long plane_rows, long kpad)
{
  dim3 grid((rows + 31) / 32, (k + 31) / 32), block(32, 8);
  oz_slice_kernel<S><<<grid, block>>>(X, ldx, rows, k, inv_scale, planes, plane_rows, kpad, 0, 0);
}

}  // namespace oz

// ---------------------------------------------------------------------------------------------------------------------
static double urand(unsigned& s) { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffffff) / 16777216.0; }

// int32 exactness of the plumbing: planes filled with known small integers, C = 0, compare every group accumulator
template <int S>
static int run_exact(int m, int n, int k)
{
  printf("exact: S %d  m %d n %d k %d  (stages %d, stage %d B, tmem cols %d)\n", S, m, n, k, oz::Cfg<S>::STAGES, oz::Cfg<S>::STAGE_BYTES,
         oz::Cfg<S>::TMEM_COLS);
  std::vector<int8_t> ha((size_t)S * m * k), hb((size_t)S * n * k);
  unsigned seed = 12345;
  for (auto& x : ha) x = (int8_t)((int)(urand(seed) * 129.0) - 64);
  for (auto& x : hb) x = (int8_t)((int)(urand(seed) * 129.0) - 64);
  int8_t *da, *db; int32_t* dbg; double* C;
  CK(cudaMalloc(&da, ha.size())); CK(cudaMalloc(&db, hb.size()));
  CK(cudaMalloc(&dbg, sizeof(int32_t) * (size_t)S * m * n)); CK(cudaMalloc(&C, sizeof(double) * (size_t)m * n));
  CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(C, 0, sizeof(double) * (size_t)m * n)); CK(cudaMemset(dbg, 0xff, sizeof(int32_t) * (size_t)S * m * n));
  oz::Args g = {};
  g.C = C; g.ldc = m; g.m = m; g.n = n; g.k = k; g.a_rows = m; g.b_rows = n; g.scale = 1.0; g.dbg = dbg;
  oz::launch<S>(da, db, g, k);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hd((size_t)S * m * n);
  CK(cudaMemcpy(hd.data(), dbg, hd.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int gi = 0; gi < S; gi++)
    for (int r = 0; r < m; r++)
      for (int c = 0; c < n; c++) {
        long long ref = 0;
        for (int i = 0; i <= gi; i++) {
          const int8_t* ap = &ha[((size_t)i * m + r) * k];
          const int8_t* bp = &hb[((size_t)(gi - i) * n + c) * k];
          for (int q = 0; q < k; q++) ref += (int)ap[q] * (int)bp[q];
        }
        const int32_t got = hd[((size_t)gi * m + r) * n + c];
        if ((long long)got != ref) { if (bad < 12) printf("  mismatch group %d row %d col %d: got %d want %lld\n", gi, r, c, got, ref); bad++; }
      }
  printf("exact: %ld mismatches of %zu\n", bad, hd.size());
  cudaFree(da); cudaFree(db); cudaFree(dbg); cudaFree(C);
  return bad != 0;
}

template <int S>
static int run_check(int m, int n, int k)
{
  std::vector<double> hA((size_t)m * k), hB((size_t)n * k), hC((size_t)m * n);
  unsigned seed = 777;
  for (auto& x : hA) x = (urand(seed) * 2 - 1) * ldexp(1.0, -(int)(urand(seed) * 12));      // 12 binades of spread, |x| <= 1
  for (auto& x : hB) x = (urand(seed) * 2 - 1) * ldexp(1.0, -(int)(urand(seed) * 12));
  for (auto& x : hC) x = urand(seed) * 2 - 1;
  double *A, *B, *C, *C2;
  int8_t *pa, *pb;
  CK(cudaMalloc(&A, hA.size() * 8)); CK(cudaMalloc(&B, hB.size() * 8)); CK(cudaMalloc(&C, hC.size() * 8)); CK(cudaMalloc(&C2, hC.size() * 8));
  CK(cudaMalloc(&pa, (size_t)S * m * k)); CK(cudaMalloc(&pb, (size_t)S * n * k));
  CK(cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(C, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(C2, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice));
  oz::slice<S>(A, m, m, k, 1.0, pa, m, k);
  oz::slice<S>(B, n, n, k, 1.0, pb, n, k);
  oz::Args g = {};
  g.C = C; g.ldc = m; g.m = m; g.n = n; g.k = k; g.a_rows = m; g.b_rows = n; g.scale = -ldexp(1.0, -2 * (oz::DIGIT_BITS - 1));
  oz::launch<S>(pa, pb, g, k);
  // the product's DMMA kernel on the same data: C2 <- C2 - A B^T
  {
    using T = GemmTileWideWS;
    CK(cudaFuncSetAttribute(gemm_nt_ws_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
    GemmArgs q = {};
    q.A = A; q.lda = m; q.B = B; q.ldb = n; q.C = C2; q.ldc = m; q.M = m; q.N = n; q.K = k; q.mt = m / T::BM; q.nt = n / T::BN;
    q.init_mode = GEMM_INIT_NEGC; q.negate_out = 1;
    gemm_nt_ws_kernel<T><<<q.mt * q.nt, T::THREADS, T::SMEM_BYTES>>>(q);
  }
  CK(cudaDeviceSynchronize());
  std::vector<double> o1(hC.size()), o2(hC.size());
  CK(cudaMemcpy(o1.data(), C, o1.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o2.data(), C2, o2.size() * 8, cudaMemcpyDeviceToHost));
  double e_oz = 0, e_dmma = 0;
  for (int r = 0; r < m; r += 7)
    for (int c = 0; c < n; c += 5) {
      long double ref = hC[r + (size_t)c * m];
      for (int q = 0; q < k; q++) ref -= (long double)hA[r + (size_t)q * m] * (long double)hB[c + (size_t)q * n];
      e_oz = fmax(e_oz, fabs((double)(o1[r + (size_t)c * m] - ref)));
      e_dmma = fmax(e_dmma, fabs((double)(o2[r + (size_t)c * m] - ref)));
    }
  const double bound = k * ldexp(1.0, -(oz::DIGIT_BITS * S - 1));      // k x (operand rounding 2^-(7S-1)), operands bounded by 1
  printf("check: S %d (%d int8 products)  m %d n %d k %d : max |C_oz - ref| %.3e  (a-priori ~ %.1e)   max |C_dmma - ref| %.3e\n", S,
         oz::Cfg<S>::PAIRS, m, n, k, e_oz, bound, e_dmma);
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(C2); cudaFree(pa); cudaFree(pb);
  return !(e_oz < 4 * bound + 1e-13);
}

template <int S>
static void run_bench(int m, int n, int k, int reps)
{
  double *A, *B, *C;
  int8_t *pa, *pb;
  CK(cudaMalloc(&A, sizeof(double) * (size_t)m * k)); CK(cudaMalloc(&B, sizeof(double) * (size_t)n * k)); CK(cudaMalloc(&C, sizeof(double) * (size_t)m * n));
  CK(cudaMalloc(&pa, (size_t)S * m * k)); CK(cudaMalloc(&pb, (size_t)S * n * k));
  CK(cudaMemset(A, 0, sizeof(double) * (size_t)m * k)); CK(cudaMemset(B, 0, sizeof(double) * (size_t)n * k)); CK(cudaMemset(C, 0, sizeof(double) * (size_t)m * n));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  oz::slice<S>(A, m, m, k, 1.0, pa, m, k);
  CK(cudaEventRecord(e0));
  oz::slice<S>(B, n, n, k, 1.0, pb, n, k);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("slice  S %d  %d x %d : %.3f ms (%.1f GB/s of FP64 read + int8 written)\n", S, n, k, ms, (8.0 + S) * n * (double)k / (ms * 1e-3) * 1e-9);
  oz::Args g = {};
  g.C = C; g.ldc = m; g.m = m; g.n = n; g.k = k; g.a_rows = m; g.b_rows = n; g.scale = -ldexp(1.0, -2 * (oz::DIGIT_BITS - 1));
  oz::launch<S>(pa, pb, g, k);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) oz::launch<S>(pa, pb, g, k);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
  const double per = ms / reps * 1e-3;
  printf("oz_gemm S %d  m %d n %d k %d : %.3f ms/launch  %.1f int8 TOP/s  = %.2f FP64-equivalent TFLOP/s\n", S, m, n, k, per * 1e3,
         2.0 * m * n * (double)k * oz::Cfg<S>::PAIRS / per * 1e-12, 2.0 * m * n * (double)k / per * 1e-12);
  {
    using T = GemmTileWideWS;
    CK(cudaFuncSetAttribute(gemm_nt_ws_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
    GemmArgs q = {};
    q.A = A; q.lda = m; q.B = B; q.ldb = n; q.C = C; q.ldc = m; q.M = m; q.N = n; q.K = k; q.mt = m / T::BM; q.nt = n / T::BN;
    q.init_mode = GEMM_INIT_NEGC; q.negate_out = 1;
    gemm_nt_ws_kernel<T><<<q.mt * q.nt, T::THREADS, T::SMEM_BYTES>>>(q);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) gemm_nt_ws_kernel<T><<<q.mt * q.nt, T::THREADS, T::SMEM_BYTES>>>(q);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dmma gemm_nt_ws          : %.3f ms/launch  %.2f TFLOP/s\n", ms / reps, 2.0 * m * n * (double)k * reps / (ms * 1e-3) * 1e-12);
  }
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(pa); cudaFree(pb);
}

int main(int argc, char** argv)
{
  const char* mode = argc > 1 ? argv[1] : "exact";
  if (!strcmp(mode, "exact")) {
    int rc = run_exact<1>(128, 64, 64);          // one tile, one chunk, one product: descriptors + TMEM read-back
    rc |= run_exact<1>(256, 128, 512);           // several tiles, stage ring wraps
    rc |= run_exact<3>(128, 64, 256);            // groups share accumulators
    rc |= run_exact<7>(256, 128, 512);
    rc |= run_exact<8>(256, 128, 512);
    return rc;
  }
  if (!strcmp(mode, "check")) {
    const int S = argc > 2 ? atoi(argv[2]) : 0;
    int rc = 0;
    if (S == 0 || S == 6) rc |= run_check<6>(512, 512, 1024);
    if (S == 0 || S == 7) rc |= run_check<7>(512, 512, 1024);
    if (S == 0 || S == 8) rc |= run_check<8>(512, 512, 1024);
    return rc;
  }
  if (!strcmp(mode, "bench")) {
    const int m = argc > 2 ? atoi(argv[2]) : 16384, n = argc > 3 ? atoi(argv[3]) : 16384, k = argc > 4 ? atoi(argv[4]) : 8192;
    const int S = argc > 5 ? atoi(argv[5]) : 7;
    if (S == 6) run_bench<6>(m, n, k, 3);
    else if (S == 8) run_bench<8>(m, n, k, 3);
    else run_bench<7>(m, n, k, 3);
    return 0;
  }
  printf("usage: %s exact | check [S] | bench [M N K S]\n", argv[0]);
  return 2;
}
