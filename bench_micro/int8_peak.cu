// int8 tensor-pipe micro-peak of sm_100a (MEASURED_PEAKS.json has no int8 entry): tcgen05.mma kind::i8 issued back to back from
// operands that never leave shared memory -- no TMA, no epilogue -- on every SM, one CTA per SM.  The instruction shapes are the
// ones oz_gemm_kernel issues (M 128, K 32, N = 64 | 128 | 192 | 256), so the table also shows what the instruction SHAPE costs:
// every instruction re-reads its A tile (4 KB) and its B tile (N x 32 B) from shared memory, and N = 64 needs 192 B/clk of
// shared-memory reads at the full tensor rate against the 128 B/clk an SM has.
//   ./int8_peak [iters]      prints int8 TOP/s per shape; the N = 256 line is the roofline denominator bench.py uses
#include "../gp_ss_ak_b200/csrc/gpss_ozaki.cuh"
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

using oz::int8_peak_kernel;

template <int N>
static double run(int iters, int sms)
{
  const int smem = 24 * 1024 + 1024 + 64;
  CK(cudaFuncSetAttribute(int8_peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int8_peak_kernel<N><<<sms, 128, smem>>>(iters / 8, cyc);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  unsigned long long hc = 0;
  for (int r = 0; r < 3; r++) {
    CK(cudaEventRecord(e0));
    int8_peak_kernel<N><<<sms, 128, smem>>>(iters, cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) { best = ms; CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost)); }
  }
  const double mmas = (double)iters * (512 / N) * 2, ops = 2.0 * 128 * N * 32 * mmas * sms;
  const double tops = ops / (best * 1e-3) * 1e-12;
  printf("int8 peak  M 128 N %3d K 32 : %8.3f ms  %7.1f int8 TOP/s  (%.1f clk per instruction on SM 0, %.0f MAC/clk/SM; %.0f B/clk of operand reads)\n", N, best, tops,
         (double)hc / mmas, 128.0 * N * 32 * mmas / (double)hc, (4096.0 + N * 32.0) * mmas / (double)hc);
  cudaFree(cyc);
  return tops;
}

int main(int argc, char** argv)
{
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<64>(iters, p.multiProcessorCount);
  run<128>(iters, p.multiProcessorCount);
  const double pk = run<256>(iters, p.multiProcessorCount);
  // sustained: ~2 s back to back (the power cap pulls the clocks down), as MEASURED_PEAKS.json does for bf16
  {
    const int smem = 24 * 1024 + 1024 + 64;
    unsigned long long* cyc; CK(cudaMalloc(&cyc, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 40;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) int8_peak_kernel<256><<<p.multiProcessorCount, 128, smem>>>(iters * 4, cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double ops = 2.0 * 128 * 256 * 32 * ((double)iters * 4 * 2 * 2) * p.multiProcessorCount * reps;
    printf("int8 peak  N 256 sustained over %.2f s : %7.1f int8 TOP/s   (burst %.1f)\n", ms * 1e-3, ops / (ms * 1e-3) * 1e-12, pk);
    printf("{\"int8_tops_burst\": %.1f, \"int8_tops_sustained\": %.1f}\n", pk, ops / (ms * 1e-3) * 1e-12);
  }
  return 0;
}
