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
//   ./ozaki_gemm tri              the product's options (triangular k-ranges, tile skipping, overwrite) against the DMMA kernel
//   ./ozaki_gemm bench [M N K S]  throughput, beside the DMMA kernel of the product on the same shape
#include "../gp_ss_ak_b200/csrc/gpss_ozaki.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

using namespace gpss;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// the kernels under test are the product's (opt-in GPSS_OZAKI path)
namespace ozh {
template <int S>
static void launch(const int8_t* pa, const int8_t* pb, oz::Args g, long kpad, cudaStream_t st = 0)
{
  static bool attr = false;
  if (!attr) { CK(oz::configure<S>()); attr = true; }
  CUtensorMap ta, tb;
  if (oz::make_plane_map(&ta, pa, (long)S * g.a_rows, kpad, oz::BM) || oz::make_plane_map(&tb, pb, (long)S * g.b_rows, kpad, oz::BN)) {
    printf("cuTensorMapEncodeTiled failed\n");
    exit(1);
  }
  oz::launch<S>(ta, tb, g, st);
}
template <int S>
static void slice(const double* X, long ldx, int rows, int k, int8_t* planes, long plane_rows, long kpad, int mask = oz::MASK_NONE)
{
  oz::slice<S>(X, ldx, 0, rows, 0, k, oz::SCALE_UNIT, mask, nullptr, planes, plane_rows, kpad, 0);
}
}  // namespace ozh

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
  g.C = C; g.ldc = m; g.m = m; g.n = n; g.k0 = 0; g.k1 = k; g.a_rows = m; g.b_rows = n; g.sign = 1.0; g.accumulate = 1; g.dbg = dbg;
  ozh::launch<S>(da, db, g, k);
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
  ozh::slice<S>(A, m, m, k, pa, m, k);
  ozh::slice<S>(B, n, n, k, pb, n, k);
  oz::Args g = {};
  g.C = C; g.ldc = m; g.m = m; g.n = n; g.k0 = 0; g.k1 = k; g.a_rows = m; g.b_rows = n; g.sign = -1.0; g.accumulate = 1;
  ozh::launch<S>(pa, pb, g, k);
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

// The options the product uses: B^-1 = U U^T as in lauum_lower (upper-triangular U, k from each tile's first row, tiles above
// the diagonal skipped, C overwritten), against the DMMA kernel with the same options.
template <int S>
static int run_tri(int n)
{
  std::vector<double> hU((size_t)n * n, 0.0);
  unsigned seed = 4242;
  for (int k = 0; k < n; k++)
    for (int r = 0; r <= k; r++) hU[r + (size_t)k * n] = (urand(seed) * 2 - 1) * ldexp(1.0, -(int)(urand(seed) * 10));
  double *U, *Q1, *Q2;
  int8_t* pu;
  CK(cudaMalloc(&U, hU.size() * 8)); CK(cudaMalloc(&Q1, hU.size() * 8)); CK(cudaMalloc(&Q2, hU.size() * 8));
  CK(cudaMalloc(&pu, (size_t)S * n * n));
  CK(cudaMemcpy(U, hU.data(), hU.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(Q1, 0, hU.size() * 8)); CK(cudaMemset(Q2, 0, hU.size() * 8));
  ozh::slice<S>(U, n, n, n, pu, n, n, oz::MASK_UPPER);
  oz::Args g = {};
  g.C = Q1; g.ldc = n; g.m = n; g.n = n; g.a_rows = n; g.b_rows = n; g.k0 = 0; g.k1 = n; g.kbeg_row = 1; g.lower_only = 1; g.sign = 1.0;
  ozh::launch<S>(pu, pu, g, n);
  {
    using T = GemmTileWideWS;
    CK(cudaFuncSetAttribute(gemm_nt_ws_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
    GemmArgs q = {};
    q.A = U; q.lda = n; q.B = U; q.ldb = n; q.C = Q2; q.ldc = n; q.M = n; q.N = n; q.K = n; q.mt = n / T::BM; q.nt = n / T::BN;
    q.lower_only = 1; q.kbeg_row = 1;
    gemm_nt_ws_kernel<T><<<q.mt * q.nt, T::THREADS, T::SMEM_BYTES>>>(q);
  }
  CK(cudaDeviceSynchronize());
  std::vector<double> o1(hU.size()), o2(hU.size());
  CK(cudaMemcpy(o1.data(), Q1, o1.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o2.data(), Q2, o2.size() * 8, cudaMemcpyDeviceToHost));
  double e = 0, mx = 0;
  for (int c = 0; c < n; c++)
    for (int r = c; r < n; r++) { e = fmax(e, fabs(o1[r + (size_t)c * n] - o2[r + (size_t)c * n])); mx = fmax(mx, fabs(o2[r + (size_t)c * n])); }
  const double bound = n * ldexp(1.0, -(oz::DIGIT_BITS * S - 1));
  printf("tri: S %d  n %d  lower triangle of U U^T: max |oz - dmma| %.3e (max |Q| %.3e, a-priori ~ %.1e)\n", S, n, e, mx, bound);
  cudaFree(U); cudaFree(Q1); cudaFree(Q2); cudaFree(pu);
  return !(e < 4 * bound + 1e-12);
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
  ozh::slice<S>(A, m, m, k, pa, m, k);
  CK(cudaEventRecord(e0));
  ozh::slice<S>(B, n, n, k, pb, n, k);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("slice  S %d  %d x %d : %.3f ms (%.1f GB/s of FP64 read + int8 written)\n", S, n, k, ms, (8.0 + S) * n * (double)k / (ms * 1e-3) * 1e-9);
  oz::Args g = {};
  g.C = C; g.ldc = m; g.m = m; g.n = n; g.k0 = 0; g.k1 = k; g.a_rows = m; g.b_rows = n; g.sign = -1.0; g.accumulate = 1;
  ozh::launch<S>(pa, pb, g, k);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) ozh::launch<S>(pa, pb, g, k);
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
  if (!strcmp(mode, "tri")) return run_tri<8>(1024) | run_tri<7>(1536);
  if (!strcmp(mode, "bench")) {
    const int m = argc > 2 ? atoi(argv[2]) : 16384, n = argc > 3 ? atoi(argv[3]) : 16384, k = argc > 4 ? atoi(argv[4]) : 8192;
    const int S = argc > 5 ? atoi(argv[5]) : 7;
    if (S == 6) run_bench<6>(m, n, k, 3);
    else if (S == 8) run_bench<8>(m, n, k, 3);
    else run_bench<7>(m, n, k, 3);
    return 0;
  }
  printf("usage: %s exact | check [S] | tri | bench [M N K S]\n", argv[0]);
  return 2;
}
