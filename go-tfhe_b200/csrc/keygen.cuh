// keygen.cuh — cloud-key generation on the device.
//
// Reference: cloudkey.NewCloudKey  cloudkey/cloudkey.go:24-31
//   genBootstrappingKey  cloudkey/cloudkey.go:122-145 : per LWE key bit s0[i]
//       trgsw.EncryptTorus      trgsw/trgsw.go:32-58    2L x trlwe.EncryptF64(0) + gadget s0[i] / Bg^(l+1)
//       trlwe.EncryptF64        trlwe/trlwe.go:28-50    A uniform, B = gaussian(0, alpha) + A * s1  (MulPoly)
//       trgsw.NewTRGSWLv1FFT    trgsw/trgsw.go:72-82    ToFourierPoly of every A and B
//   genKeySwitchingKey   cloudkey/cloudkey.go:88-120 : row (base*t*i + base*j + k), k >= 1, =
//       tlwe.EncryptF64(k * s1[i] / 2^((j+1)*basebit), alpha, s0)   tlwe/tlwe.go:36-52 ; k = 0 rows stay zero
//   utils.F64ToTorus / GaussianF64  utils/utils.go:11-49  (frac(d) * 2^32 truncated; mu and noise converted separately)
//
// The reference draws from unseeded math/rand, so there is nothing to match bit for bit: the outputs here follow the
// same distributions (uniform masks, N(0, alpha^2) noise truncated to the torus the same way) from a counter-based
// generator, one independent stream per ciphertext, so a key depends only on (secret key, seed).  Both kernels write
// the REFERENCE layouts (FourierPoly groups of 4 re + 4 im; rows of n+1 words), i.e. exactly the CloudKey fields a Go
// caller holds; the engine then ingests them through the same tfhe_ctx_load_cloudkey_device path as an uploaded key.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blind_rotate.cuh"

namespace tfhe {

__host__ __device__ __forceinline__ uint64_t kg_mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// stream key of ciphertext `id` in domain `dom` (0 = key-switching rows, 1 = TRLWE rows of the bootstrapping key)
__host__ __device__ __forceinline__ uint64_t kg_stream(uint64_t seed, uint64_t dom, uint64_t id) {
  return kg_mix(seed ^ kg_mix(2 * id + dom + 0x632BE59BD9B4E019ull));
}
__device__ __forceinline__ uint64_t kg_u64(uint64_t key, uint64_t idx) { return kg_mix(key + 0xD1342543DE82EF95ull * (idx + 1)); }
__device__ __forceinline__ double kg_unit(uint64_t w) { return (double)((w >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
// standard normal, Box-Muller on words (2 idx, 2 idx + 1) of the stream
__device__ __forceinline__ double kg_gauss(uint64_t key, uint64_t idx) {
  const double u = kg_unit(kg_u64(key, 2 * idx)), v = kg_unit(kg_u64(key, 2 * idx + 1));
  return sqrt(-2.0 * log(u)) * cospi(2.0 * v);
}
// utils.F64ToTorus (utils/utils.go:11-14): Torus(int64(math.Mod(d, 1.0) * 2^32))
__device__ __forceinline__ uint32_t kg_to_torus(double d) {
  const double f = fmod(d, 1.0) * 4294967296.0;
  return (uint32_t)(unsigned long long)(long long)f;
}

// Key-switching key, reference layout [N*t*base][n+1].  grid = rows, block = 128.
// Words [n + 1 ...) of the mask stream are the noise words, so mask and noise never share a counter.
__global__ void __launch_bounds__(128) keygen_ksk_kernel(uint32_t* __restrict__ ksk, const uint32_t* __restrict__ s0,
                                                         const uint32_t* __restrict__ s1, int n, int basebit, int t,
                                                         double alpha, uint64_t seed) {
  __shared__ uint32_t red[4];
  const size_t row = blockIdx.x;
  uint32_t* dst = ksk + row * (size_t)(n + 1);
  const int k = (int)(row & ((1u << basebit) - 1u));
  if (k == 0) {  // never read by IdentityKeySwitching; the reference leaves NewTLWELv0() zeros there
    for (int w = threadIdx.x; w <= n; w += blockDim.x) dst[w] = 0u;
    return;
  }
  const size_t ij = row >> basebit;
  const int j = (int)(ij % t);
  const size_t i = ij / t;
  const uint64_t key = kg_stream(seed, 0, row);
  uint32_t dot = 0;
  for (int w = threadIdx.x; w < n; w += blockDim.x) {
    const uint32_t a = (uint32_t)(kg_u64(key, w) >> 32);
    dst[w] = a;
    dot += a * s0[w];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t inner = red[0] + red[1] + red[2] + red[3];
    const double mu = ((double)k * (double)s1[i]) / (double)(1ull << ((j + 1) * basebit));
    const double z = kg_gauss(key, (uint64_t)n + 1);
    dst[n] = inner + kg_to_torus(mu) + kg_to_torus(z * alpha);
  }
}

struct KeygenBskArgs {
  double* bsk_fft;           // [n][2L][2][N] doubles, reference FourierPoly layout
  const uint32_t* s0;        // [n]  LWE key bits
  const uint32_t* s1;        // [N]  ring key bits
  const double2* tw_tab;
  double alpha;
  uint64_t seed;
  int L, bgbit;
  Tw4 tw0;
};

// One block per TRLWE row of the bootstrapping key: grid = n * 2L, block = N/16 threads.
template <int LOGN>
__global__ void __launch_bounds__((1 << (LOGN - 4)), 2) keygen_bsk_kernel(const KeygenBskArgs A) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* ex = reinterpret_cast<double2*>(smem_raw);
  const int tau = threadIdx.x;
  const size_t rowid = blockIdx.x;          // = i * 2L + r
  const int r = (int)(rowid % (2 * A.L));
  const size_t i = rowid / (2 * A.L);
  Fft<LOGN - 1, false> fft;
  fft.init(ex, A.tw_tab, tau);
  __syncthreads();
  const uint64_t key = kg_stream(A.seed, 1, rowid);

  // mask polynomial (words [0, N) of the stream) and the ring key, folded as ToFourierPoly does (int32 view)
  uint32_t are[8], aim[8];
  double2 x[8], y[8];
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
    are[a] = (uint32_t)(kg_u64(key, j) >> 32);
    aim[a] = (uint32_t)(kg_u64(key, j + M) >> 32);
    x[a] = make_double2((double)(int32_t)are[a], (double)(int32_t)aim[a]);
    y[a] = make_double2((double)(int32_t)A.s1[j], (double)(int32_t)A.s1[j + M]);
  }
  fft.forward(x, A.tw0);
  fft.forward(y, A.tw0);
  const double inv_m = 1.0 / (double)M;
#pragma unroll
  for (int e = 0; e < 8; e++) {  // poly/fourier_ops.go:138-161 (MulPoly), scaled for the inverse
    const double re = (x[e].x * y[e].x - x[e].y * y[e].y) * inv_m;
    const double im = (x[e].x * y[e].y + x[e].y * y[e].x) * inv_m;
    y[e] = make_double2(re, im);
  }
  fft.inverse(y, A.tw0);
  // B = gaussian(0, alpha) + A * s1 ; then the gadget term s0[i] / Bg^(l+1) on A[0] (rows < L) or B[0] (rows >= L)
  const uint32_t bit = A.s0[i];
  const int l = (r < A.L) ? r : r - A.L;
  const uint32_t gadget = bit * (1u << (32 - (l + 1) * A.bgbit));
  uint32_t bre[8], bim[8];
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
    bre[a] = to_torus<false>(y[a].x) + kg_to_torus(0.0) + kg_to_torus(kg_gauss(key, (uint64_t)N + j) * A.alpha);
    bim[a] = to_torus<false>(y[a].y) + kg_to_torus(0.0) + kg_to_torus(kg_gauss(key, (uint64_t)N + j + M) * A.alpha);
    if (j == 0) {
      if (r < A.L) are[a] += gadget;
      else bre[a] += gadget;
    }
  }
  // ToFourierPoly of A and B, stored in the reference layout: complex k = 8 tau + e has re at (k/4)*8 + k%4, im 4 further
  double* oa = A.bsk_fft + (rowid * 2 + 0) * (size_t)N;
  double* ob = oa + N;
#pragma unroll
  for (int a = 0; a < 8; a++) {
    x[a] = make_double2((double)(int32_t)are[a], (double)(int32_t)aim[a]);
    y[a] = make_double2((double)(int32_t)bre[a], (double)(int32_t)bim[a]);
  }
  fft.forward(x, A.tw0);
  fft.forward(y, A.tw0);
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int k = 8 * tau + e;
    const int p = (k >> 2) * 8 + (k & 3);
    oa[p] = x[e].x; oa[p + 4] = x[e].y;
    ob[p] = y[e].x; ob[p + 4] = y[e].y;
  }
}

}  // namespace tfhe
