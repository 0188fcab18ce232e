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
// same distributions (uniform masks, N(0, alpha^2) noise truncated to the torus the same way).  Randomness is ChaCha20
// (chacha.h) under a 256-bit key taken from the OS entropy source by the host (or expanded from a caller's seed for
// reproducible tests), one stream per ciphertext and SEPARATE domains for the public masks and the secret noise.  Both
// kernels write the REFERENCE layouts (FourierPoly groups of 4 re + 4 im; rows of n+1 words), i.e. exactly the CloudKey
// fields a Go caller holds; the engine then ingests them through the same tfhe_ctx_load_cloudkey_device path as an
// uploaded key.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blind_rotate.cuh"
#include "chacha.h"

namespace tfhe {

// standard normal from four stream words (Box-Muller)
__device__ __forceinline__ double kg_gauss4(const uint32_t* w) {
  const double u = rng_unit(w[0], w[1]), v = rng_unit(w[2], w[3]);
  return sqrt(-2.0 * log(u)) * cospi(2.0 * v);
}
// utils.F64ToTorus (utils/utils.go:11-14): Torus(int64(math.Mod(d, 1.0) * 2^32))
__device__ __forceinline__ uint32_t kg_to_torus(double d) {
  const double f = fmod(d, 1.0) * 4294967296.0;
  return (uint32_t)(unsigned long long)(long long)f;
}

// Key-switching key, reference layout [N*t*base][n+1].  grid = rows, block = 128.
__global__ void __launch_bounds__(128) keygen_ksk_kernel(uint32_t* __restrict__ ksk, const uint32_t* __restrict__ s0,
                                                         const uint32_t* __restrict__ s1, int n, int basebit, int t,
                                                         double alpha, const RngKey key) {
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
  uint32_t dot = 0;
  for (int blk = threadIdx.x; blk * 16 < n; blk += blockDim.x) {  // mask: block blk of stream (KSK_MASK, row)
    uint32_t w[16];
    chacha20_block(key, (uint32_t)blk, RNG_KSK_MASK, (uint32_t)row, (uint32_t)(row >> 32), w);
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int p = blk * 16 + q;
      if (p < n) { dst[p] = w[q]; dot += w[q] * s0[p]; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t inner = red[0] + red[1] + red[2] + red[3];
    const double mu = ((double)k * (double)s1[i]) / (double)(1ull << ((j + 1) * basebit));
    uint32_t w[16];
    chacha20_block(key, 0, RNG_KSK_NOISE, (uint32_t)row, (uint32_t)(row >> 32), w);
    dst[n] = inner + kg_to_torus(mu) + kg_to_torus(kg_gauss4(w) * alpha);
  }
}

struct KeygenBskArgs {
  double* bsk_fft;           // [n][2L][2][N] doubles, reference FourierPoly layout
  const uint32_t* s0;        // [n]  LWE key bits
  const uint32_t* s1;        // [N]  ring key bits
  const double2* tw_tab;
  double alpha;
  RngKey key;
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
  uint32_t* rnd = reinterpret_cast<uint32_t*>(smem_raw + (size_t)br_nbuf(LOGN) * TFHE_BR_EXW * M * 16);  // [N] mask words
  Fft<LOGN - 1, false> fft;
  fft.init(ex, A.tw_tab, tau);
  {  // mask polynomial: the N words of stream (BSK_MASK, row); thread tau draws block tau (T = N/16 threads)
    uint32_t w[16];
    chacha20_block(A.key, (uint32_t)tau, RNG_BSK_MASK, (uint32_t)rowid, (uint32_t)(rowid >> 32), w);
#pragma unroll
    for (int q = 0; q < 16; q++) rnd[16 * tau + q] = w[q];
  }
  __syncthreads();

  // mask and ring key, folded as ToFourierPoly does (int32 view)
  uint32_t are[8], aim[8];
  double2 x[8], y[8];
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
    are[a] = rnd[j];
    aim[a] = rnd[j + M];
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
  // noise: stream (BSK_NOISE, row); thread tau draws blocks 4 tau .. 4 tau + 3 = four words for each of its 16 coefficients
  uint32_t bre[8], bim[8];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint32_t w[16];
    chacha20_block(A.key, (uint32_t)(4 * tau + q), RNG_BSK_NOISE, (uint32_t)rowid, (uint32_t)(rowid >> 32), w);
#pragma unroll
    for (int h = 0; h < 4; h++) {
      const int a = (4 * q + h) & 7;
      const double z = kg_gauss4(w + 4 * h) * A.alpha;
      if (4 * q + h < 8) bre[a] = to_torus<false>(y[a].x) + kg_to_torus(0.0) + kg_to_torus(z);
      else bim[a] = to_torus<false>(y[a].y) + kg_to_torus(0.0) + kg_to_torus(z);
    }
  }
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
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
