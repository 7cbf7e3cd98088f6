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

// shared memory: A tile 128 rows x 64 B (SWIZZLE_64B layout, 8 KB) + B tile 256 rows x 64 B (16 KB), filled with pseudo-random bytes
// (all-zero operands would flatter the clocks: tensor power is data dependent)
template <int N>
__global__ void __launch_bounds__(128, 1) int8_peak_kernel(int iters, unsigned long long* cycles)
{
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  unsigned s = 1234567u + blockIdx.x * 7919u + threadIdx.x;
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) { s = s * 1664525u + 1013904223u; smem[i] = (uint8_t)((int)((s >> 16) % 129u) - 64); }
  if (threadIdx.x == 0) { gpss::mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(gpss::smem_u32(slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // generic-proxy writes of the operands -> visible to the tensor core
  oz::tc_fence_before();
  __syncthreads();
  oz::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = gpss::smem_u32(smem), sb = sa + 8 * 1024;
    const long long t0 = clock64();
    constexpr int ACC = 512 / N;                  // independent accumulators of N columns (N = 192: 2)
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int q = 0; q < ACC; q++) {
#pragma unroll
        for (int ks = 0; ks < 2; ks++)
          oz::mma_i8(tmem + (uint32_t)(q * N), oz::smem_desc_k<64>(sa + ks * 32), oz::smem_desc_k<64>(sb + ks * 32), oz::idesc_i8(N), it > 0 || ks > 0);
      }
    }
    oz::tc_commit(bar);
    gpss::mbar_wait(bar, 0);
    if (blockIdx.x == 0) *cycles = (unsigned long long)(clock64() - t0);
  }
  oz::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    oz::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u) : "memory");
  }
}

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
