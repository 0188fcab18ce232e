// blind_rotate_w16.cuh — warp-per-gate blind rotation for N = 1024 (M = 512-point folded transforms), sm_100a.
//
// Same arithmetic and the same reference functions replaced as blind_rotate.cuh (evaluator/evaluator.go:50-135,
// poly/decomposer.go:55-66, poly/fourier_transform.go, poly/fourier_ops.go:167-191, poly/buffer_methods.go:133-164,
// trlwe/trlwe_ops.go:10-21), remapped to cut the shared-memory traffic that bounds the block-per-gate kernel
// (profiles/r01_experiments.md: LSU data pipe 72 %, FP64 pipe 55 %):
//
//   * ONE WARP per gate, 16 complex points per thread.  A transform is radix-16 (registers) -> one swizzled
//     shared-memory exchange -> radix-16 (registers) -> a half exchange with the neighbouring lane by warp shuffle
//     -> the last radix-2 stage.  One 8 KiB exchange per transform instead of two, and no block barrier anywhere
//     in the step loop (only __syncwarp).
//   * The two spectrum accumulators (2 x 512 complex doubles per gate = 16 KiB) do not fit in registers next to 16
//     points per thread, so they live in TENSOR MEMORY: each thread owns 128 TMEM columns of its lane and
//     read-modify-writes them with tcgen05.ld / tcgen05.st (SASS LDTM / STTM) in 16-column chunks during the
//     multiply-accumulate.  TMEM is used purely as per-thread scratch; no MMA is involved.
//   * A CTA is four independent warps (four gates), one per TMEM lane quadrant.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blind_rotate.cuh"

#ifndef TFHE_TM_PF_L1
#define TFHE_TM_PF_L1 0   // d > 0: prefetch.global.L1 of the key rows d digits ahead (block-per-gate TMEM kernel)
#endif
#ifndef TFHE_TM_EARLY_BK
#define TFHE_TM_EARLY_BK 0
#endif
#ifndef TFHE_TM_BK_SPLIT
#define TFHE_TM_BK_SPLIT 0
#endif
#ifndef TFHE_W16_PREFETCH
#define TFHE_W16_PREFETCH 1
#endif

namespace tfhe {

struct Tw8 { double2 s[8]; };  // radix-16 block (m,i): S(m,i), S(2m,2i), S(4m,4i), S(4m,4i+2), S(8m,8i+{0,2,4,6})

struct BrW16Args {
  const uint32_t* ct_in;
  const uint32_t* testvec;
  const uint32_t* luts;
  long long nluts;
  long long count;
  const double2* bsk;      // [n][2L][2][16 slots][32 lanes]; slot s<8: position 16*lane+2s, s>=8: 16*lane+2(s-8)+1
  const double2* tw1;      // pass-1 twiddles [8][16]: entry j of block b at tw1[j*16 + b]
  const double2* twl;      // last-stage twiddles [4][32]: S(256, 8*lane + 2j) at twl[j*32 + lane]
  uint32_t* out;
  int n;
  uint32_t offset;
  int out_mode;
  Tw8 tw0;                 // pass-0 twiddles (kernel constants)
};

// ---- TMEM primitives (32 lanes x 32-bit columns per warp; every thread touches only its own lane) -----------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void pack4(const double2 (&v)[4], uint32_t (&r)[16]) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    r[4 * k + 0] = (uint32_t)__double2loint(v[k].x);
    r[4 * k + 1] = (uint32_t)__double2hiint(v[k].x);
    r[4 * k + 2] = (uint32_t)__double2loint(v[k].y);
    r[4 * k + 3] = (uint32_t)__double2hiint(v[k].y);
  }
}
__device__ __forceinline__ void unpack4(const uint32_t (&r)[16], double2 (&v)[4]) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    v[k].x = __hiloint2double((int)r[4 * k + 1], (int)r[4 * k + 0]);
    v[k].y = __hiloint2double((int)r[4 * k + 3], (int)r[4 * k + 2]);
  }
}

__device__ __forceinline__ void pack2(const double2 (&v)[2], uint32_t (&r)[8]) {
#pragma unroll
  for (int k = 0; k < 2; k++) {
    r[4 * k + 0] = (uint32_t)__double2loint(v[k].x);
    r[4 * k + 1] = (uint32_t)__double2hiint(v[k].x);
    r[4 * k + 2] = (uint32_t)__double2loint(v[k].y);
    r[4 * k + 3] = (uint32_t)__double2hiint(v[k].y);
  }
}
__device__ __forceinline__ void unpack2(const uint32_t (&r)[8], double2 (&v)[2]) {
#pragma unroll
  for (int k = 0; k < 2; k++) {
    v[k].x = __hiloint2double((int)r[4 * k + 1], (int)r[4 * k + 0]);
    v[k].y = __hiloint2double((int)r[4 * k + 3], (int)r[4 * k + 2]);
  }
}

// ---- radix-16 register blocks: stage A on (lo[k], hi[k]) then two radix-8 blocks ---------------------------------
__device__ __forceinline__ void radix16_fwd(double2 (&lo)[8], double2 (&hi)[8], const double2 (&s)[8]) {
#pragma unroll
  for (int k = 0; k < 8; k++) bf_fwd(lo[k], hi[k], s[0].x, s[0].y);
  radix8_fwd<3>(lo, s[1], s[2], s[4], s[5]);
  radix8_fwd<3>(hi, make_double2(s[1].y, -s[1].x), s[3], s[6], s[7]);  // block 2i+1: -i * S(2m,2i)
}
__device__ __forceinline__ void radix16_inv(double2 (&lo)[8], double2 (&hi)[8], const double2 (&s)[8]) {
  radix8_inv<3>(lo, s[1], s[2], s[4], s[5]);
  radix8_inv<3>(hi, make_double2(s[1].y, -s[1].x), s[3], s[6], s[7]);
#pragma unroll
  for (int k = 0; k < 8; k++) bf_inv(lo[k], hi[k], s[0].x, s[0].y);
}

// swizzle for the one exchange: writes hit 32 consecutive 16-byte slots, reads hit {32b + u + 2a'} for four
// consecutive b per quarter warp; xor-ing bits 1-2 with b makes both conflict-free.
__device__ __forceinline__ int swz16(int p) { return p ^ (((p >> 5) & 3) << 1); }

struct FftW16 {
  double2* ex;   // this warp's exchange buffer [512]
  int lane;
  double2 tl[4]; // last-stage twiddles S(256, 8 lane + {0,2,4,6})
  const double2* tw1;

  __device__ __forceinline__ void init(double2* ex_, const double2* tw1_, const double2* twl, int lane_) {
    ex = ex_; lane = lane_; tw1 = tw1_;
#pragma unroll
    for (int j = 0; j < 4; j++) tl[j] = __ldg(twl + j * 32 + lane);
  }
  __device__ __forceinline__ void load_tw1(double2 (&s)[8]) const {
    const int b = lane >> 1;
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = __ldg(tw1 + j * 16 + b);
  }
  // lanes 2b and 2b+1 trade halves so that each ends with 8 complete (even, odd) pairs: (lo[k], hi[k])
  __device__ __forceinline__ void swap_halves(double2 (&lo)[8], double2 (&hi)[8]) const {
    const bool u = lane & 1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double sx = u ? lo[k].x : hi[k].x, sy = u ? lo[k].y : hi[k].y;
      const double rx = __shfl_xor_sync(0xffffffffu, sx, 1), ry = __shfl_xor_sync(0xffffffffu, sy, 1);
      lo[k].x = u ? rx : lo[k].x; lo[k].y = u ? ry : lo[k].y;
      hi[k].x = u ? hi[k].x : rx; hi[k].y = u ? hi[k].y : ry;
    }
  }
  // in: lo[k] = z[lane + 32k], hi[k] = z[lane + 256 + 32k];  out: lo[k] = Y[16 lane + 2k], hi[k] = Y[16 lane + 2k + 1]
  __device__ __forceinline__ void forward(double2 (&lo)[8], double2 (&hi)[8], const Tw8& tw0) {
    radix16_fwd(lo, hi, tw0.s);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      ex[swz16(32 * k + lane)] = lo[k];
      ex[swz16(32 * (k + 8) + lane)] = hi[k];
    }
    __syncwarp();
    const int rb = 32 * (lane >> 1) + (lane & 1);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      lo[k] = ex[swz16(rb + 2 * k)];
      hi[k] = ex[swz16(rb + 2 * (k + 8))];
    }
    double2 s[8];
    load_tw1(s);
    radix16_fwd(lo, hi, s);
    swap_halves(lo, hi);
  }
  // the last radix-2 stage of forward(); split off so that the caller can start key / accumulator loads before it
  __device__ __forceinline__ void forward_last(double2 (&lo)[8], double2 (&hi)[8]) const {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double2 w = tl[k >> 1];
      if (k & 1) bf_fwd(lo[k], hi[k], w.y, -w.x); else bf_fwd(lo[k], hi[k], w.x, w.y);
    }
  }
  __device__ __forceinline__ void inverse(double2 (&lo)[8], double2 (&hi)[8], const Tw8& tw0) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double2 w = tl[k >> 1];
      if (k & 1) bf_inv(lo[k], hi[k], w.y, -w.x); else bf_inv(lo[k], hi[k], w.x, w.y);
    }
    swap_halves(lo, hi);
    double2 s[8];
    load_tw1(s);
    radix16_inv(lo, hi, s);
    const int wb = 32 * (lane >> 1) + (lane & 1);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      ex[swz16(wb + 2 * k)] = lo[k];
      ex[swz16(wb + 2 * (k + 8))] = hi[k];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      lo[k] = ex[swz16(32 * k + lane)];
      hi[k] = ex[swz16(32 * (k + 8) + lane)];
    }
    radix16_inv(lo, hi, tw0.s);
  }
};

__host__ __device__ constexpr size_t br_w16_warp_smem(int n) { return (size_t)8192 /*acc*/ + 8192 /*exchange*/ + (size_t)((n * 2 + 15) / 16 * 16) /*abar*/; }
constexpr size_t br_w16_smem_bytes(int n) { return 4 * br_w16_warp_smem(n); }

template <int L, int BGBIT, bool SMALL, int MINB>
__global__ void __launch_bounds__(128, MINB) blind_rotate_w16_kernel(const BrW16Args A) {
  constexpr int N = 1024, M = 512, LOGN = 10;
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = A.n;
  unsigned char* wsm = smem_raw + (size_t)warp * br_w16_warp_smem(n);
  uint32_t* acc = reinterpret_cast<uint32_t*>(wsm);                 // [2][N]
  double2* ex = reinterpret_cast<double2*>(wsm + 8192);             // [M]
  unsigned short* abar = reinterpret_cast<unsigned short*>(wsm + 16384);

  // 128 TMEM columns for the CTA: every thread owns columns [0,64) = A accumulator, [64,128) = B accumulator of its lane
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = s_tmem_base + ((uint32_t)(warp * 32) << 16);

  // Warps past the end of the batch recompute the last gate and skip the store: keeps every warp on one
  // uniform path (no divergence scopes around the shuffles and TMEM ops).
  const long long gid = (long long)blockIdx.x * 4 + warp;
  const bool active = gid < A.count;
  const long long g = active ? gid : A.count - 1;
  {
    const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);
    for (int i = lane; i < n; i += 32) abar[i] = (unsigned short)((ct[i] + (1u << (30 - LOGN))) >> (31 - LOGN));
    const unsigned long long bb = (unsigned long long)ct[n] + (1ull << (30 - LOGN));
    const int btil = (int)((2 * N - (int)(bb >> (31 - LOGN))) & (2 * N - 1));
    const uint32_t* __restrict__ tv = A.luts ? A.luts + (A.nluts == 1 ? 0 : g) * (2 * N) : A.testvec;
    for (int j = lane; j < N; j += 32) {
      const int idx = (j - btil) & (2 * N - 1);
      const uint32_t va = tv[idx & (N - 1)], vb = tv[N + (idx & (N - 1))];
      acc[j] = (idx & N) ? ~va : va;
      acc[N + j] = (idx & N) ? ~vb : vb;
    }
    FftW16 fft;
    fft.init(ex, A.tw1, A.twl, lane);
    __syncwarp();

    const size_t row_stride = (size_t)2 * L * 2 * M;
    for (int i = 0; i < n; i++) {
      const int at = abar[i];
      if (at == 0) continue;  // exact no-op
      const double2* __restrict__ bk = A.bsk + row_stride * i + lane;
#pragma unroll 1
      for (int poly = 0; poly < 2; poly++) {
      // (X^at * P - P + offset) at this thread's 32 coefficients, shared by the L decomposition levels
      uint32_t d[32];
      {
        const uint32_t* P = acc + poly * N;
        const int ib = (lane - at) & (2 * N - 1);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          // fold: real part from coefficient j, imaginary part from j + 512, for j = lane + 32k and j + 256
          const int j0 = lane + 32 * k, j1 = j0 + 256;
          d[4 * k + 0] = rot_read<N>(P, ib + 32 * k) - P[j0] + A.offset;
          d[4 * k + 1] = rot_read<N>(P, ib + 32 * k + M) - P[j0 + M] + A.offset;
          d[4 * k + 2] = rot_read<N>(P, ib + 32 * k + 256) - P[j1] + A.offset;
          d[4 * k + 3] = rot_read<N>(P, ib + 32 * k + 256 + M) - P[j1 + M] + A.offset;
        }
      }
#pragma unroll 1
      for (int lvl = 0; lvl < L; lvl++) {
        const int r = poly * L + lvl;
        const int sh = 32 - (lvl + 1) * BGBIT;
#if TFHE_W16_PREFETCH
        {  // this digit's key row-set (A and B spectra, 16 KiB = 128 lines) -> L1, a whole transform ahead of its use
          const char* row = reinterpret_cast<const char*>(A.bsk + row_stride * i + (size_t)r * 2 * M);
#pragma unroll
          for (int q = 0; q < 4; q++) asm volatile("prefetch.global.L1 [%0];" ::"l"(row + (lane + 32 * q) * 128));
        }
#endif
        double2 lo[8], hi[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          lo[k].x = field_to_double((d[4 * k + 0] >> sh) & MASK, BIAS);
          lo[k].y = field_to_double((d[4 * k + 1] >> sh) & MASK, BIAS);
          hi[k].x = field_to_double((d[4 * k + 2] >> sh) & MASK, BIAS);
          hi[k].y = field_to_double((d[4 * k + 3] >> sh) & MASK, BIAS);
        }
        fft.forward(lo, hi, A.tw0);
        // Multiply-accumulate into the TMEM-resident spectra, four complex points (16 columns) at a time, software
        // pipelined: the key values and accumulator columns of chunk c+1 are in flight while chunk c is computed,
        // and chunk 0 is requested before the last butterfly stage of the transform.
        const double2* __restrict__ rowA = bk + (size_t)(r * 2 + 0) * M;
        const double2* __restrict__ rowB = rowA + M;
        double2 ka[2][2], kb[2][2];
        uint32_t ra[2][8], rb[2][8];
#pragma unroll
        for (int q = 0; q < 2; q++) { ka[0][q] = __ldg(rowA + q * 32); kb[0][q] = __ldg(rowB + q * 32); }
        if (r > 0) { tmem_ld8(tacc, ra[0]); tmem_ld8(tacc + 64, rb[0]); }
        fft.forward_last(lo, hi);
#pragma unroll
        for (int c = 0; c < 8; c++) {  // chunk c: slots 2c, 2c+1 (lo for c<4, hi for c>=4)
          const int cur = c & 1, nxt = cur ^ 1;
          if (c < 7) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
              ka[nxt][q] = __ldg(rowA + (2 * (c + 1) + q) * 32);
              kb[nxt][q] = __ldg(rowB + (2 * (c + 1) + q) * 32);
            }
          }
          double2 aA[2], aB[2];
          if (r > 0) {
            tmem_wait_ld();
            unpack2(ra[cur], aA);
            unpack2(rb[cur], aB);
            if (c < 7) { tmem_ld8(tacc + 8 * (c + 1), ra[nxt]); tmem_ld8(tacc + 64 + 8 * (c + 1), rb[nxt]); }
          } else {
#pragma unroll
            for (int q = 0; q < 2; q++) { aA[q] = make_double2(0.0, 0.0); aB[q] = make_double2(0.0, 0.0); }
          }
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const double2 xv = (c < 4) ? lo[2 * c + q] : hi[2 * (c - 4) + q];
            aA[q].x = fma(xv.x, ka[cur][q].x, aA[q].x);
            aA[q].x = fma(-xv.y, ka[cur][q].y, aA[q].x);
            aA[q].y = fma(xv.x, ka[cur][q].y, aA[q].y);
            aA[q].y = fma(xv.y, ka[cur][q].x, aA[q].y);
            aB[q].x = fma(xv.x, kb[cur][q].x, aB[q].x);
            aB[q].x = fma(-xv.y, kb[cur][q].y, aB[q].x);
            aB[q].y = fma(xv.x, kb[cur][q].y, aB[q].y);
            aB[q].y = fma(xv.y, kb[cur][q].x, aB[q].y);
          }
          uint32_t sa[8], sb[8];
          pack2(aA, sa);
          pack2(aB, sb);
          tmem_st8(tacc + 8 * c, sa);
          tmem_st8(tacc + 64 + 8 * c, sb);
        }
        tmem_wait_st();
      }
      }
      // two inverse transforms out of TMEM, result added into the accumulator polynomials
#pragma unroll 1
      for (int poly = 0; poly < 2; poly++) {
        double2 lo[8], hi[8];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          uint32_t ra[16];
          double2 v[4];
          tmem_ld16(tacc + 64 * poly + 16 * c, ra);
          tmem_wait_ld();
          unpack4(ra, v);
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (c < 2) lo[4 * c + q] = v[q]; else hi[4 * (c - 2) + q] = v[q];
          }
        }
        fft.inverse(lo, hi, A.tw0);
        uint32_t* P = acc + poly * N;
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int j0 = lane + 32 * k, j1 = j0 + 256;
          P[j0] += to_torus<SMALL>(lo[k].x);
          P[j0 + M] += to_torus<SMALL>(lo[k].y);
          P[j1] += to_torus<SMALL>(hi[k].x);
          P[j1 + M] += to_torus<SMALL>(hi[k].y);
        }
      }
      __syncwarp();
    }

    if (active) {
      if (A.out_mode == 0) {
        uint32_t* o = A.out + g * (2 * N);
        for (int j = lane; j < 2 * N; j += 32) o[j] = acc[j];
      } else {
        uint32_t* o = A.out + g * (N + 1);
        for (int j = lane; j < N; j += 32) o[j] = (j == 0) ? acc[0] : ~acc[N - j];
        if (lane == 0) o[N] = acc[N];
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(s_tmem_base) : "memory");
}

// =============================================================================================
// Block-per-gate kernel with the spectrum accumulators in TMEM (variant "tmem").
// Same mapping as blind_rotate_kernel (T = N/16 threads, 8 points per thread, two exchanges per transform), but the
// 2 x 8 complex accumulators per thread (64 registers) live in 64 TMEM columns of the thread's lane and are
// read-modify-written in 16-column chunks by the multiply-accumulate.  That takes the kernel from 255 to <= 168
// registers, i.e. from 4 to 6 resident blocks per SM, to overlap the shared-memory and FP64 pipes better.
// =============================================================================================
// TFHE_BR_TM_MAXNREG: experiment knob — cap the register count directly (e.g. 200 -> 5 blocks per SM) instead of
// letting ptxas derive it from the minimum-blocks bound (which only ever picks 255, 168 or 128).
#ifdef TFHE_BR_TM_MAXNREG
#define TFHE_BR_TM_BOUNDS(T, MINB) __maxnreg__(TFHE_BR_TM_MAXNREG)
#else
#define TFHE_BR_TM_BOUNDS(T, MINB) TFHE_BR_BOUNDS(T, MINB)
#endif
template <int LOGN, int L, int BGBIT, bool SMALL, int MINB>
__global__ void TFHE_BR_TM_BOUNDS((1 << (LOGN - 4)), MINB) blind_rotate_tm_kernel(const BrArgs A) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  static_assert(T >= 32 && T <= 128, "one TMEM lane per thread");
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_tmem_base;
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                    // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);               // [2][EXW][M]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + 16 * br_nbuf(LOGN) * TFHE_BR_EXW * M);
  const int tau = threadIdx.x;
  const long long g = blockIdx.x;
  const int n = A.n;
  const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);

  if (tau < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tau; i < n; i += T) abar[i] = (unsigned short)((ct[i] + (1u << (30 - LOGN))) >> (31 - LOGN));
  const unsigned long long bb = (unsigned long long)ct[n] + (1ull << (30 - LOGN));
  const int btil = (int)((2 * N - (int)(bb >> (31 - LOGN))) & (2 * N - 1));
  const uint32_t* __restrict__ tv = A.luts ? A.luts + (A.nluts == 1 ? 0 : g) * (2 * N) : A.testvec;
  for (int j = tau; j < N; j += T) {
    const int idx = (j - btil) & (2 * N - 1);
    const uint32_t va = tv[idx & (N - 1)], vb = tv[N + (idx & (N - 1))];
    acc[j] = (idx & N) ? ~va : va;
    acc[N + j] = (idx & N) ? ~vb : vb;
  }
  Fft<LOGN - 1, false> fft;
  fft.init(ex, A.tw_tab, tau);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this thread's lane: warp w of the block owns TMEM lanes [32w, 32w+32); columns [0,32) = A, [32,64) = B accumulator
  const uint32_t tacc = s_tmem_base + ((uint32_t)(tau & ~31) << 16);

  const size_t row_stride = (size_t)2 * L * 2 * M;
  for (int i = 0; i < n; i++) {
    const int at = abar[i];
    if (at == 0) continue;
    const double2* __restrict__ bk = A.bsk + row_stride * i + tau;
#pragma unroll 1
    for (int poly = 0; poly < 2; poly++) {
      const uint32_t* P = acc + poly * N;
      uint32_t dre[8], dim[8];
      int ib = (tau - at) & (2 * N - 1);
      asm volatile("" : "+r"(ib));  // keep the 32 rotated indices from being hoisted out of the loop and spilled
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        dre[a] = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
        dim[a] = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
      }
      int lvl0 = 0;
#if TFHE_TM_PAIR
      // two levels at a time: one block barrier and one burst of shared-memory traffic per exchange for both transforms
      auto mac_tm = [&](const double2 (&x)[8], int r) {
        const double2* __restrict__ rowA = bk + (size_t)(r * 2 + 0) * M;
        const double2* __restrict__ rowB = rowA + M;
        double2 ka[8], kb[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { ka[e] = __ldg(rowA + e * T); kb[e] = __ldg(rowB + e * T); }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double2 aA[4], aB[4];
          uint32_t ra[16], rb[16];
          if (r > 0) {
            tmem_ld16(tacc + 16 * h, ra);
            tmem_ld16(tacc + 32 + 16 * h, rb);
            tmem_wait_ld();
            unpack4(ra, aA);
            unpack4(rb, aB);
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) { aA[q] = make_double2(0.0, 0.0); aB[q] = make_double2(0.0, 0.0); }
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int e = 4 * h + q;
            aA[q].x = fma(x[e].x, ka[e].x, aA[q].x);
            aA[q].x = fma(-x[e].y, ka[e].y, aA[q].x);
            aA[q].y = fma(x[e].x, ka[e].y, aA[q].y);
            aA[q].y = fma(x[e].y, ka[e].x, aA[q].y);
            aB[q].x = fma(x[e].x, kb[e].x, aB[q].x);
            aB[q].x = fma(-x[e].y, kb[e].y, aB[q].x);
            aB[q].y = fma(x[e].x, kb[e].y, aB[q].y);
            aB[q].y = fma(x[e].y, kb[e].x, aB[q].y);
          }
          pack4(aA, ra);
          pack4(aB, rb);
          tmem_st16(tacc + 16 * h, ra);
          tmem_st16(tacc + 32 + 16 * h, rb);
        }
        tmem_wait_st();
      };
#pragma unroll 1
      for (; lvl0 + 1 < L; lvl0 += 2) {
        double2 x[8], y[8];
        const int sh0 = 32 - (lvl0 + 1) * BGBIT, sh1 = 32 - (lvl0 + 2) * BGBIT;
#pragma unroll
        for (int a = 0; a < 8; a++) {
          x[a].x = field_to_double((dre[a] >> sh0) & MASK, BIAS);
          x[a].y = field_to_double((dim[a] >> sh0) & MASK, BIAS);
          y[a].x = field_to_double((dre[a] >> sh1) & MASK, BIAS);
          y[a].y = field_to_double((dim[a] >> sh1) & MASK, BIAS);
        }
        fft.forward2(x, y, A.tw0);
        mac_tm(x, poly * L + lvl0);
        mac_tm(y, poly * L + lvl0 + 1);
      }
#endif
#pragma unroll 1
      for (int lvl = lvl0; lvl < L; lvl++) {
        const int r = poly * L + lvl;
        const int sh = 32 - (lvl + 1) * BGBIT;
        double2 x[8];
#pragma unroll
        for (int a = 0; a < 8; a++) {
          x[a].x = field_to_double((dre[a] >> sh) & MASK, BIAS);
          x[a].y = field_to_double((dim[a] >> sh) & MASK, BIAS);
        }
        const double2* __restrict__ rowA = bk + (size_t)(r * 2 + 0) * M;
        const double2* __restrict__ rowB = rowA + M;
#if TFHE_TM_PF_L1
        {  // L1 prefetch of the NEXT digit's rows (contiguous into the next step's row-set): no registers held in flight
          const char* pf = reinterpret_cast<const char*>(A.bsk + row_stride * i + (size_t)(r + TFHE_TM_PF_L1) * 2 * M) + tau * 128;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + T * 128));
        }
#endif
#if TFHE_TM_EARLY_BK
        // all 16 key values requested BEFORE the last register pass: the L2 latency hides behind 72 FMAs
        fft.forward_head(x, A.tw0);
        double2 ka[8], kb[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { ka[e] = __ldg(rowA + e * T); kb[e] = __ldg(rowB + e * T); }
        uint32_t ra0[16], rb0[16];
        if (r > 0) { tmem_ld16(tacc, ra0); tmem_ld16(tacc + 32, rb0); }
        fft.forward_tail(x, A.tw0);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double2 aA[4], aB[4];
          uint32_t ra[16], rb[16];
          if (r > 0) {
            if (h == 1) { tmem_ld16(tacc + 16, ra); tmem_ld16(tacc + 48, rb); }
            tmem_wait_ld();
            if (h == 0) { unpack4(ra0, aA); unpack4(rb0, aB); } else { unpack4(ra, aA); unpack4(rb, aB); }
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) { aA[q] = make_double2(0.0, 0.0); aB[q] = make_double2(0.0, 0.0); }
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int e = 4 * h + q;
            aA[q].x = fma(x[e].x, ka[e].x, aA[q].x);
            aA[q].x = fma(-x[e].y, ka[e].y, aA[q].x);
            aA[q].y = fma(x[e].x, ka[e].y, aA[q].y);
            aA[q].y = fma(x[e].y, ka[e].x, aA[q].y);
            aB[q].x = fma(x[e].x, kb[e].x, aB[q].x);
            aB[q].x = fma(-x[e].y, kb[e].y, aB[q].x);
            aB[q].y = fma(x[e].x, kb[e].y, aB[q].y);
            aB[q].y = fma(x[e].y, kb[e].x, aB[q].y);
          }
          pack4(aA, ra);
          pack4(aB, rb);
          tmem_st16(tacc + 16 * h, ra);
          tmem_st16(tacc + 32 + 16 * h, rb);
        }
#elif TFHE_TM_BK_SPLIT
        fft.forward(x, A.tw0);
        // key values and accumulator columns one half (4 points) at a time, next half in flight: 32 + 32 registers
        double2 ka[2][4], kb[2][4];
#pragma unroll
        for (int q = 0; q < 4; q++) { ka[0][q] = __ldg(rowA + q * T); kb[0][q] = __ldg(rowB + q * T); }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          if (h == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) { ka[1][q] = __ldg(rowA + (4 + q) * T); kb[1][q] = __ldg(rowB + (4 + q) * T); }
          }
          double2 aA[4], aB[4];
          uint32_t ra[16], rb[16];
          if (r > 0) {
            tmem_ld16(tacc + 16 * h, ra);
            tmem_ld16(tacc + 32 + 16 * h, rb);
            tmem_wait_ld();
            unpack4(ra, aA);
            unpack4(rb, aB);
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) { aA[q] = make_double2(0.0, 0.0); aB[q] = make_double2(0.0, 0.0); }
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int e = 4 * h + q;
            aA[q].x = fma(x[e].x, ka[h][q].x, aA[q].x);
            aA[q].x = fma(-x[e].y, ka[h][q].y, aA[q].x);
            aA[q].y = fma(x[e].x, ka[h][q].y, aA[q].y);
            aA[q].y = fma(x[e].y, ka[h][q].x, aA[q].y);
            aB[q].x = fma(x[e].x, kb[h][q].x, aB[q].x);
            aB[q].x = fma(-x[e].y, kb[h][q].y, aB[q].x);
            aB[q].y = fma(x[e].x, kb[h][q].y, aB[q].y);
            aB[q].y = fma(x[e].y, kb[h][q].x, aB[q].y);
          }
          pack4(aA, ra);
          pack4(aB, rb);
          tmem_st16(tacc + 16 * h, ra);
          tmem_st16(tacc + 32 + 16 * h, rb);
        }
#else
        fft.forward(x, A.tw0);
        double2 ka[8], kb[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { ka[e] = __ldg(rowA + e * T); kb[e] = __ldg(rowB + e * T); }  // all 16 in flight
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double2 aA[4], aB[4];
          uint32_t ra[16], rb[16];
          if (r > 0) {
            tmem_ld16(tacc + 16 * h, ra);
            tmem_ld16(tacc + 32 + 16 * h, rb);
            tmem_wait_ld();
            unpack4(ra, aA);
            unpack4(rb, aB);
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) { aA[q] = make_double2(0.0, 0.0); aB[q] = make_double2(0.0, 0.0); }
          }
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const int e = 4 * h + q;
            aA[q].x = fma(x[e].x, ka[e].x, aA[q].x);
            aA[q].x = fma(-x[e].y, ka[e].y, aA[q].x);
            aA[q].y = fma(x[e].x, ka[e].y, aA[q].y);
            aA[q].y = fma(x[e].y, ka[e].x, aA[q].y);
            aB[q].x = fma(x[e].x, kb[e].x, aB[q].x);
            aB[q].x = fma(-x[e].y, kb[e].y, aB[q].x);
            aB[q].y = fma(x[e].x, kb[e].y, aB[q].y);
            aB[q].y = fma(x[e].y, kb[e].x, aB[q].y);
          }
          pack4(aA, ra);
          pack4(aB, rb);
          tmem_st16(tacc + 16 * h, ra);
          tmem_st16(tacc + 32 + 16 * h, rb);
        }
#endif
        tmem_wait_st();
      }
    }
#if TFHE_TM_PAIR
    {
      double2 x[8], y[8];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t ra[16], rb[16];
        double2 v[4], w[4];
        tmem_ld16(tacc + 16 * h, ra);
        tmem_ld16(tacc + 32 + 16 * h, rb);
        tmem_wait_ld();
        unpack4(ra, v);
        unpack4(rb, w);
#pragma unroll
        for (int q = 0; q < 4; q++) { x[4 * h + q] = v[q]; y[4 * h + q] = w[q]; }
      }
      fft.inverse2(x, y, A.tw0);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        acc[j] += to_torus<SMALL>(x[a].x);
        acc[j + M] += to_torus<SMALL>(x[a].y);
        acc[N + j] += to_torus<SMALL>(y[a].x);
        acc[N + j + M] += to_torus<SMALL>(y[a].y);
      }
    }
#else
#pragma unroll 1
    for (int poly = 0; poly < 2; poly++) {
      double2 x[8];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint32_t ra[16];
        double2 v[4];
        tmem_ld16(tacc + 32 * poly + 16 * h, ra);
        tmem_wait_ld();
        unpack4(ra, v);
#pragma unroll
        for (int q = 0; q < 4; q++) x[4 * h + q] = v[q];
      }
      fft.inverse(x, A.tw0);
      uint32_t* P = acc + poly * N;
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        P[j] += to_torus<SMALL>(x[a].x);
        P[j + M] += to_torus<SMALL>(x[a].y);
      }
    }
#endif
    __syncthreads();
  }

  if (A.out_mode == 0) {
    uint32_t* o = A.out + g * (2 * N);
    for (int j = tau; j < 2 * N; j += T) o[j] = acc[j];
  } else {
    uint32_t* o = A.out + g * (N + 1);
    for (int j = tau; j < N; j += T) o[j] = (j == 0) ? acc[0] : ~acc[N - j];
    if (tau == 0) o[N] = acc[N];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tau < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(s_tmem_base) : "memory");
}

// Bootstrapping key, reference FourierPoly layout -> the warp-per-gate layout above (scaled by 1/M).
__global__ void bsk_repack_w16_kernel(const double* __restrict__ src, double2* __restrict__ dst) {
  constexpr int N = 1024, M = 512;
  const size_t poly = blockIdx.x;
  const double scale = 1.0 / (double)M;
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    const double re = src[poly * N + (k >> 2) * 8 + (k & 3)];
    const double im = src[poly * N + (k >> 2) * 8 + 4 + (k & 3)];
    const int lane = k >> 4, w = k & 15;          // position k = 16 lane + w
    const int slot = (w >> 1) + ((w & 1) ? 8 : 0);  // even positions -> slots 0..7, odd -> 8..15
    dst[poly * M + slot * 32 + lane] = make_double2(re * scale, im * scale);
  }
}

}  // namespace tfhe
