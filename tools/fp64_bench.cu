// DFMA pipe microbenchmark: throughput (DFMA per clock per SM) as a function of warps per scheduler and of the number
// of independent dependency chains per thread.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_bench fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(64) chain(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) acc[j] = threadIdx.x + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int j = 0; j < ILP; j++) acc[j] = fma(acc[j], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += acc[j];
  if (s == 12345.678) out[0] = s;
}
template <int ILP>
void run(int blocks_per_sm, int sms, double* d) {
  const int iters = 4096;
  size_t smem = 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // limit residency with dynamic shared memory so that exactly blocks_per_sm 64-thread blocks share an SM
  smem = (200 * 1024) / blocks_per_sm;
  cudaFuncSetAttribute(chain<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  chain<ILP><<<sms * blocks_per_sm, 64, smem>>>(d, 16, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  chain<ILP><<<sms * blocks_per_sm, 64, smem>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double dfma_per_sm = (double)iters * 8 * ILP * 64 * blocks_per_sm;
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("ILP %2d  warps/SMSP %.1f  : %.2f thread-DFMA/clk/SM  (%.2f ms)  => latency-bound chain step %.1f clk\n", ILP,
         blocks_per_sm * 2 / 4.0, dfma_per_sm / cycles, ms, cycles / (iters * 8.0));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d; cudaMalloc(&d, 8);
  for (int b : {2, 4, 6, 8, 16}) {
    run<1>(b, sms, d); run<2>(b, sms, d); run<4>(b, sms, d); run<8>(b, sms, d); run<16>(b, sms, d);
  }
  return 0;
}
