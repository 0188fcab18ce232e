// Is the FP64 tensor path (mma.sync m8n8k4 f64, SASS DMMA) a pipe of its own on B200, i.e. can it run next to DFMA?
// Times (a) a DFMA stream, (b) a DMMA stream, (c) both interleaved in one warp, (d) half the warps each.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(64) k(double* out, const double* in, int iters) {
  double acc[8], c[8], a = in[threadIdx.x], b = in[64 + threadIdx.x], x = in[128 + threadIdx.x];
#pragma unroll
  for (int j = 0; j < 8; j++) { acc[j] = in[threadIdx.x + j]; c[j] = in[200 + threadIdx.x + j]; }
  const bool dm = (MODE == 1) || (MODE == 2) || (MODE == 3 && (threadIdx.x >> 5) == 1);
  const bool df = (MODE == 0) || (MODE == 2) || (MODE == 3 && (threadIdx.x >> 5) == 0);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (df) {
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = fma(acc[j], 2.0, -x);   // 8 DFMA (immediate form: full pipe rate)
      }
      if (dm) {
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(c[2 * j], c[2 * j + 1], a, b);  // 4 DMMA = 4 * 256 FMA per warp
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j] + c[j];
  if (s == 12345.678) out[0] = s;
}
template <int MODE>
float run(int bps, int sms, double* d, const double* in, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t smem = (200 * 1024) / bps;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<MODE><<<sms * bps, 64, smem>>>(d, in, 16);
  cudaEventRecord(e0);
  k<MODE><<<sms * bps, 64, smem>>>(d, in, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 8 * 1024); cudaMemset(in, 0, 8 * 1024);
  const int iters = 2048;
  for (int bps : {2, 4}) {
    const float t0 = run<0>(bps, sms, d, in, iters), t1 = run<1>(bps, sms, d, in, iters), t2 = run<2>(bps, sms, d, in, iters), t3 = run<3>(bps, sms, d, in, iters);
    const double cyc = clk * 1e3 * 1e-3;
    const double dfma = (double)iters * 4 * 8 * 64 * bps, dmma_fma = (double)iters * 4 * 4 * 256 * 2 * bps;
    printf("warps/SMSP %.1f: DFMA only %.3f ms (%.1f FMA/clk/SM) | DMMA only %.3f ms (%.1f FMA/clk/SM) | both in every warp %.3f ms (sum of parts %.3f) | split by warp %.3f ms (max of half-parts %.3f)\n",
           bps * 2 / 4.0, t0, dfma / (t0 * cyc), t1, dmma_fma / (t1 * cyc), t2, t0 + t1, t3, (t0 > t1 ? t0 : t1) / 2);
  }
  return 0;
}
