// Standalone throughput harness for the FP64 DMMA GEMM-NT kernel (gp_ss_ak_b200/csrc/gpss_gemm.cuh).
//   ./gemm_bench [M N K reps]     prints TFLOP/s per tile configuration; used under ncu for source-level stalls.
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

template <class T>
static double run_ws(const char* name, double* A, double* B, double* C, int M, int N, int K, int reps, int negc)
{
  CK(cudaFuncSetAttribute(gemm_nt_ws_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
  GemmArgs g = {};
  g.A = A; g.lda = M; g.B = B; g.ldb = N; g.C = C; g.ldc = M; g.M = M; g.N = N; g.K = K;
  g.mt = M / T::BM; g.nt = N / T::BN;
  if (negc) { g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  gemm_nt_ws_kernel<T><<<g.mt * g.nt, T::THREADS, T::SMEM_BYTES>>>(g);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) gemm_nt_ws_kernel<T><<<g.mt * g.nt, T::THREADS, T::SMEM_BYTES>>>(g);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double tf = 2.0 * M * N * (double)K * reps / (ms * 1e-3) * 1e-12;
  printf("%-28s M %d N %d K %d negc %d : %.3f ms/launch  %.2f TFLOP/s\n", name, M, N, K, negc, ms / reps, tf);
  return tf;
}

static double max_abs_diff(const double* dA, const double* dB, size_t n)
{
  std::vector<double> a(n), b(n);
  CK(cudaMemcpy(a.data(), dA, n * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), dB, n * sizeof(double), cudaMemcpyDeviceToHost));
  double m = 0;
  for (size_t i = 0; i < n; i++) { double d = a[i] - b[i]; if (d < 0) d = -d; if (d > m) m = d; }
  return m;
}

template <class T>
static double run(const char* name, double* A, double* B, double* C, int M, int N, int K, int reps, int negc)
{
  CK(cudaFuncSetAttribute(gemm_nt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM_BYTES));
  GemmArgs g = {};
  g.A = A; g.lda = M; g.B = B; g.ldb = N; g.C = C; g.ldc = M; g.M = M; g.N = N; g.K = K;
  g.mt = M / T::BM; g.nt = N / T::BN;
  if (negc) { g.init_mode = GEMM_INIT_NEGC; g.negate_out = 1; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  gemm_nt_kernel<T><<<g.mt * g.nt, T::THREADS, T::SMEM_BYTES>>>(g);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; r++) gemm_nt_kernel<T><<<g.mt * g.nt, T::THREADS, T::SMEM_BYTES>>>(g);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double tf = 2.0 * M * N * (double)K * reps / (ms * 1e-3) * 1e-12;
  printf("%-28s M %d N %d K %d negc %d : %.3f ms/launch  %.2f TFLOP/s\n", name, M, N, K, negc, ms / reps, tf);
  return tf;
}

int main(int argc, char** argv)
{
  int M = argc > 1 ? atoi(argv[1]) : 16384, N = argc > 2 ? atoi(argv[2]) : 16384, K = argc > 3 ? atoi(argv[3]) : 4096;
  int reps = argc > 4 ? atoi(argv[4]) : 3;
  double *A, *B, *C;
  CK(cudaMalloc(&A, sizeof(double) * (size_t)M * K));
  CK(cudaMalloc(&B, sizeof(double) * (size_t)N * K));
  CK(cudaMalloc(&C, sizeof(double) * (size_t)M * N));
  fill_kernel<<<1024, 256>>>(A, (size_t)M * K, 1);
  fill_kernel<<<1024, 256>>>(B, (size_t)N * K, 2);
  fill_kernel<<<1024, 256>>>(C, (size_t)M * N, 3);
  CK(cudaDeviceSynchronize());
  run<GemmTileWide>("wide 128x64 4w 2cta", A, B, C, M, N, K, reps, 0);
  {
    double* C2;
    CK(cudaMalloc(&C2, sizeof(double) * (size_t)M * N));
    run_ws<GemmTileWideWS>("wide WS bulk+mbarrier", A, B, C2, M, N, K, reps, 0);
    printf("   max |C_ws - C_ref| = %.3e (bitwise-equal expected: same k order)\n", max_abs_diff(C, C2, (size_t)M * N));
    run_ws<GemmTileWS<128, 64, 2, 2, 2, 3>>("wide WS 3 stages", A, B, C2, M, N, K, reps, 0);
    if (K >= 2048) run_ws<GemmTileWideWS>("wide WS k=512 negc", A, B, C2, M, N, 512, reps * 4, 1);
    CK(cudaFree(C2));
  }
  run<GemmTile<128, 128, 2, 4, 1>>("legacy 128x128 8w 1cta", A, B, C, M, N, K, reps, 0);
#ifdef GEMM_EXTRA_VARIANTS
  GEMM_EXTRA_VARIANTS
#endif
  if (K >= 2048) run<GemmTileWide>("wide k=512", A, B, C, M, N, 512, reps * 4, 1);
  return 0;
}
