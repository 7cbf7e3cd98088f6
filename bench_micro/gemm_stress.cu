// Stress test: warp-specialised kernel vs the legacy cp.async kernel on the launch shapes of the potrf driver
// (in-place panel solves, small k, NEGC, lower_only), repeated; results must be bitwise identical.
#include "../gp_ss_ak_b200/csrc/gpss_gemm.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace gpss;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_kernel(double* p, size_t n, unsigned seed)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u + seed;
    x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
    p[i] = (double)(x & 0xffff) / 65536.0 - 0.5;
  }
}
template <class T> static void launch_ws(GemmArgs g, cudaStream_t st = 0)
{
  g.mt = g.M / T::BM; g.nt = g.N / T::BN;
  gemm_nt_ws_kernel<T><<<g.mt * g.nt, T::THREADS, T::SMEM_BYTES, st>>>(g);
}
static void launch_legacy(GemmArgs g, cudaStream_t st = 0)
{
  using T = GemmTileWide;
  g.mt = g.M / T::BM; g.nt = g.N / T::BN;
  gemm_nt_kernel<T><<<g.mt * g.nt, T::THREADS, T::SMEM_BYTES, st>>>(g);
}
static size_t count_diff(const double* a, const double* b, size_t n)
{
  std::vector<double> ha(n), hb(n);
  CK(cudaMemcpy(ha.data(), a, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), b, n * 8, cudaMemcpyDeviceToHost));
  size_t c = 0; long first = -1;
  for (size_t i = 0; i < n; i++) if (ha[i] != hb[i]) { if (first < 0) first = (long)i; c++; }
  if (c) printf("      first diff at linear %ld\n", first);
  return c;
}
int main(int argc, char** argv)
{
  using T = GemmTileWideWS;
  CK(cudaFuncSetAttribute(gemm_nt_ws_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm_nt_kernel<GemmTileWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmTileWide::SMEM_BYTES));
  const int ld = 6144, reps = argc > 1 ? atoi(argv[1]) : 50;
  const size_t nn = (size_t)ld * ld;
  double *A0, *A1, *A2, *W;
  CK(cudaMalloc(&A0, nn * 8)); CK(cudaMalloc(&A1, nn * 8)); CK(cudaMalloc(&A2, nn * 8)); CK(cudaMalloc(&W, 128 * 128 * 8));
  fill_kernel<<<1024, 256>>>(A0, nn, 1);
  fill_kernel<<<16, 256>>>(W, 128 * 128, 2);
  CK(cudaDeviceSynchronize());
  struct Shape { const char* name; int kind; };
  Shape shapes[] = {{"trsm g1 (in place, k=128)", 0}, {"trsm g2 (in place, k=64)", 1}, {"inner update (negc, lower, k=128)", 2},
                    {"deep update (negc, lower, k=512)", 3}, {"plain k=1024", 4}};
  for (auto& sh : shapes) {
    size_t bad_launches = 0;
    for (int r = 0; r < reps; r++) {
      CK(cudaMemcpy(A1, A0, nn * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(A2, A0, nn * 8, cudaMemcpyDeviceToDevice));
      const int k = 128 * (1 + r % 8), m = ld - k - 128;
      for (int v = 0; v < 2; v++) {
        double* A = v ? A2 : A1;
        double* A21 = A + (long)k * ld + (k + 128);
        GemmArgs g = {};
        if (sh.kind == 0) { g.A = A21; g.lda = ld; g.B = W + 64; g.ldb = 128; g.C = A21 + 64 * ld; g.ldc = ld; g.M = m; g.N = 64; g.K = 128; }
        if (sh.kind == 1) { g.A = A21; g.lda = ld; g.B = W; g.ldb = 128; g.C = A21; g.ldc = ld; g.M = m; g.N = 64; g.K = 64; }
        if (sh.kind == 2) { g.A = A21; g.lda = ld; g.B = A21; g.ldb = ld; g.C = A + (long)(k + 128) * ld + (k + 128); g.ldc = ld; g.M = m; g.N = 384; g.K = 128;
                            g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = k + 128; g.gcol0 = k + 128; }
        if (sh.kind == 3) { const int T0 = 1024; const double* Lp = A + T0; g.A = Lp; g.lda = ld; g.B = Lp; g.ldb = ld; g.C = A + (long)T0 * ld + T0; g.ldc = ld;
                            g.M = ld - T0; g.N = 512; g.K = 512; g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; g.lower_only = 1; g.grow0 = T0; g.gcol0 = T0; }
        if (sh.kind == 4) { g.A = A; g.lda = ld; g.B = A + 2048; g.ldb = ld; g.C = A + (long)2048 * ld; g.ldc = ld; g.M = 2048; g.N = 2048; g.K = 1024; }
        if (v) launch_ws<T>(g); else launch_legacy(g);
      }
      CK(cudaDeviceSynchronize());
      size_t d = count_diff(A1, A2, nn);
      if (d) { bad_launches++; printf("   %s rep %d: %zu differing entries\n", sh.name, r, d); }
    }
    printf("%-40s : %zu / %d launches differ\n", sh.name, bad_launches, reps);
  }
  // ---- concurrency test: two update-shaped GEMMs on two streams, disjoint outputs, overlapping inputs ----
  {
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    auto conc = [&](const char* name, auto tag) {
      using TC = decltype(tag);
      CK(cudaFuncSetAttribute(gemm_nt_ws_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::SMEM_BYTES));
      size_t bad = 0;
      for (int r = 0; r < reps; r++) {
        CK(cudaMemcpy(A1, A0, nn * 8, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(A2, A0, nn * 8, cudaMemcpyDeviceToDevice));
        const int T0 = 2048, T1 = 2560;
        for (int v = 0; v < 2; v++) {
          double* A = v ? A2 : A1;
          GemmArgs x = {}; { const double* Lp = A + (long)(T0 - 512) * ld + T0; x.A = Lp; x.lda = ld; x.B = Lp; x.ldb = ld; x.C = A + (long)T0 * ld + T0; x.ldc = ld;
            x.M = ld - T0; x.N = 512; x.K = 512; x.init_mode = GEMM_INIT_NEGC; x.negate_out = 1; x.lower_only = 1; x.grow0 = T0; x.gcol0 = T0; }
          GemmArgs y = {}; { const double* Lp = A + T1; y.A = Lp; y.lda = ld; y.B = Lp; y.ldb = ld; y.C = A + (long)T1 * ld + T1; y.ldc = ld;
            y.M = ld - T1; y.N = 512; y.K = T0; y.init_mode = GEMM_INIT_NEGC; y.negate_out = 1; y.lower_only = 1; y.grow0 = T1; y.gcol0 = T1; }
          if (!v) { launch_legacy(x, s1); CK(cudaStreamSynchronize(s1)); launch_legacy(y, s2); CK(cudaStreamSynchronize(s2)); }
          else { launch_ws<TC>(y, s2); launch_ws<TC>(x, s1); }
          CK(cudaDeviceSynchronize());
        }
        size_t d = count_diff(A1, A2, nn);
        if (d) bad++;
      }
      printf("concurrent %-46s: %zu / %d differ\n", name, bad, reps);
    };
    conc("product kernel (128x64, 4 stages, 2 CTAs/SM)", GemmTileWideWS());
  }
  // ---- chain test: the potrf_panel launch sequence (without the diagonal-block kernel), back to back, no host sync ----
  for (int mode = 0; mode < 4; mode++) {
    size_t bad = 0;
    for (int r = 0; r < reps / 4 + 1; r++) {
      CK(cudaMemcpy(A1, A0, nn * 8, cudaMemcpyDeviceToDevice));
      CK(cudaMemcpy(A2, A0, nn * 8, cudaMemcpyDeviceToDevice));
      for (int v = 0; v < 2; v++) {
        double* A = v ? A2 : A1;
        for (int K0 = 0; K0 < 2048; K0 += 512)
          for (int k = K0; k < K0 + 512; k += 128) {
            const int m = ld - k - 128;
            double* A21 = A + (long)k * ld + (k + 128);
            GemmArgs g1 = {}; g1.A = A21; g1.lda = ld; g1.B = W + 64; g1.ldb = 128; g1.C = A21 + 64 * ld; g1.ldc = ld; g1.M = m; g1.N = 64; g1.K = 128;
            GemmArgs g2 = {}; g2.A = A21; g2.lda = ld; g2.B = W; g2.ldb = 128; g2.C = A21; g2.ldc = ld; g2.M = m; g2.N = 64; g2.K = 64;
            const int ncols = K0 + 512 - (k + 128);
            GemmArgs g3 = {}; g3.A = A21; g3.lda = ld; g3.B = A21; g3.ldb = ld; g3.C = A + (long)(k + 128) * ld + (k + 128); g3.ldc = ld; g3.M = m; g3.N = ncols; g3.K = 128;
            g3.init_mode = GEMM_INIT_NEGC; g3.negate_out = 1; g3.lower_only = 1; g3.grow0 = k + 128; g3.gcol0 = k + 128;
            if (mode == 0 || mode == 1) { if (v) launch_ws<T>(g1); else launch_legacy(g1); }
            if (mode == 0 || mode == 2) { if (v) launch_ws<T>(g2); else launch_legacy(g2); }
            if ((mode == 0 || mode == 3) && ncols > 0) { if (v) launch_ws<T>(g3); else launch_legacy(g3); }
          }
        CK(cudaDeviceSynchronize());
      }
      size_t d = count_diff(A1, A2, nn);
      if (d) { bad++; printf("   chain mode %d rep %d: %zu differing entries\n", mode, r, d); }
    }
    printf("chain mode %d (0 = g1+g2+update, 1 = g1 only, 2 = g2 only, 3 = update only): %zu runs differ\n", mode, bad);
  }
  return 0;
}
