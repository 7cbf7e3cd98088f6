// FP64 micro-peaks on B200 (sm_100a): DFMA vs DMMA register-only loops.
// Establishes the roofline denominator for the Cholesky/SYRK path (MEASURED_PEAKS.json has no FP64 entry).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b)
{
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: 256 FMA per warp instruction
template<int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b)
{
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = 0; c1[i] = 0; }
  double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(fa), "d"(fb));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k8: 1024 FMA per warp instruction (A: 4 regs, B: 2 regs, C: 4 regs)
template<int NACC>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a, double b)
{
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0]=c[i][1]=c[i][2]=c[i][3]=0; }
  double fa = a + threadIdx.x * 1e-9, fb = b;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(fa), "d"(fa), "d"(fa), "d"(fa), "d"(fb), "d"(fb));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F>
float time_it(F f, int reps)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("device %s sms %d cc %d.%d\n", p.name, nsm, p.major, p.minor);
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 256));
  const int iters = 20000;
  for (int bps = 1; bps <= 8; bps *= 2) {
    int grid = nsm * bps;
    float ms = time_it([&]{ dfma_kernel<16><<<grid,256>>>(out, iters, 0.999999, 1e-7); }, 5);
    double fl = 2.0 * grid * 256.0 * 16 * iters;
    printf("DFMA      ilp16 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
  }
  for (int bps = 1; bps <= 4; bps *= 2) {
    int grid = nsm * bps;
    float ms = time_it([&]{ dmma884_kernel<16><<<grid,256>>>(out, iters, 0.999999, 1e-7); }, 5);
    double fl = 2.0 * grid * 8.0 * 256 * 16 * iters;
    printf("DMMA m8n8k4   x16 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    ms = time_it([&]{ dmma884_kernel<32><<<grid,256>>>(out, iters, 0.999999, 1e-7); }, 5);
    fl = 2.0 * grid * 8.0 * 256 * 32 * iters;
    printf("DMMA m8n8k4   x32 blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    ms = time_it([&]{ dmma1688_kernel<8><<<grid,256>>>(out, iters, 0.999999, 1e-7); }, 5);
    fl = 2.0 * grid * 8.0 * 1024 * 8 * iters;
    printf("DMMA m16n8k8  x8  blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
  }
  CK(cudaDeviceSynchronize());
  // sustained: 3 s of DMMA to see clocks under power cap
  {
    int grid = nsm * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int launches = 0;
    for (int r = 0; r < 40; r++) { dmma884_kernel<32><<<grid,256>>>(out, iters * 4, 0.999999, 1e-7); launches++; }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * grid * 8.0 * 256 * 32 * (iters * 4.0) * launches;
    printf("DMMA m8n8k4 sustained: %.1f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
  }
  return 0;
}
