// DFMA operand-bandwidth microbenchmark: does a DFMA with three distinct 64-bit register sources issue as fast as one
// whose multiplicands come from the operand-reuse cache / an immediate?   ILP = 8 chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(64) chain(double* out, const double* in, int iters) {
  double acc[8], x[8], y[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { acc[j] = in[threadIdx.x + j]; x[j] = in[64 + threadIdx.x + j]; y[j] = in[128 + threadIdx.x + j]; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (MODE == 0) acc[j] = fma(acc[j], x[0], y[0]);            // two shared multiplicands (reuse cache)
        if (MODE == 1) acc[j] = fma(x[j], y[(j + u) & 7], acc[j]);  // three distinct register pairs
        if (MODE == 2) acc[j] = fma(x[j], y[0], acc[j]);            // one shared
        if (MODE == 3) acc[j] = fma(acc[j], 2.0, -x[j]);            // immediate + two pairs
        if (MODE == 4) acc[j] = acc[j] + x[j];                      // DADD two pairs
        if (MODE == 5) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[j]) : "d"(x[j]), "d"(y[(j + 3) & 7]));  // order pinned: no reuse possible
        if (MODE == 6) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[j]) : "d"(x[0]), "d"(y[(j + 3) & 7]));  // order pinned: one operand shared
      }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j];
  if (s == 12345.678) out[0] = s;
}
template <int MODE>
void run(int blocks_per_sm, int sms, double* d, const double* in) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t smem = (200 * 1024) / blocks_per_sm;
  cudaFuncSetAttribute(chain<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  chain<MODE><<<sms * blocks_per_sm, 64, smem>>>(d, in, 16);
  cudaEventRecord(e0);
  chain<MODE><<<sms * blocks_per_sm, 64, smem>>>(d, in, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double n = (double)iters * 64 * 64 * blocks_per_sm;
  printf("mode %d  warps/SMSP %.1f : %.2f thread-DFMA/clk/SM\n", MODE, blocks_per_sm * 2 / 4.0, n / (ms * 1e-3 * clk * 1e3));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 8 * 512); cudaMemset(in, 0, 8 * 512);
  for (int b : {2, 4, 8}) { run<0>(b, sms, d, in); run<1>(b, sms, d, in); run<2>(b, sms, d, in); run<3>(b, sms, d, in); run<4>(b, sms, d, in); run<5>(b, sms, d, in); run<6>(b, sms, d, in); }
  return 0;
}
