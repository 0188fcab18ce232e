// fp64_probe.cuh — measured FP64 roof of the device the engine runs on (bench.py's roofline_fp64.peak).
// MEASURED_PEAKS.json carries an HBM and a bf16 tensor figure but no FP64 one, and the blind rotation is bound by the FP64
// and shared-memory pipes, not by HBM.  The probe runs independent DFMA chains whose multiplicands come from the
// operand-reuse cache (the friendliest case: 8 chains per thread, 8 warps per scheduler) and reports 2 flops per DFMA.
#pragma once
#include <cuda_runtime.h>

namespace tfhe {

__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, const double* in, int iters) {
  double acc[8];
  const double x = in[0], y = in[1];
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = in[2 + ((threadIdx.x + j) & 15)];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = fma(acc[j], x, y);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += acc[j];
  if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace tfhe
