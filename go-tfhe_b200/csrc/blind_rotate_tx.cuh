// blind_rotate_tx.cuh — block-per-gate blind rotation for N = 1024 whose SECOND transform exchange goes through
// tensor memory instead of shared memory (variant "tmex").
//
// Same arithmetic, same pass structure (three radix-8 register passes on M = 512 points, 64 threads per gate) and the
// same reference functions replaced as blind_rotate.cuh.  What changes is how the 8x8 transposes between pass 1 and
// pass 2 move: the measured bound of the shared-memory kernel is the LSU data pipe (profiles/r01_experiments.md), and
// that exchange is half of its traffic.  Here every warp
//   1. stores its 8 points per thread to TMEM with tcgen05.st.32x32b (thread -> its own lane),
//   2. reads them back with tcgen05.ld.16x256b, whose fragment layout hands thread t the doubles (4n + t%4) of lanes
//      t/4 + 8g: two register-index bits move into the lane index and two lane bits move out (probe:
//      profiles/r01_tmem_ld_shapes_probe.txt),
//   3. finishes the third bit with one half-swap with lane^4 by warp shuffle,
// and lands exactly in the layout the last pass (and the key layout) already use.  The first exchange (across the two
// warps of the block) stays in shared memory.  Thread <-> (block b, offset u) assignment of the middle pass is chosen
// so that the bits TMEM moves are the right ones: lane = (u2 u1 b1 b0 u0), warp = b2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blind_rotate.cuh"
#include "blind_rotate_w16.cuh"

#ifndef TFHE_TX_UNROLL_POLY
#define TFHE_TX_UNROLL_POLY 2
#endif
#ifndef TFHE_TX_UNROLL_LVL
#define TFHE_TX_UNROLL_LVL 1
#endif

namespace tfhe {

__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// swizzle of the one shared-memory exchange under the (u2 u1 b1 b0 u0) lane assignment: writes are 64 consecutive
// slots per register; reads of a quarter warp hit {64 b + u0 + 8 a'} for four b — xor bits 1-2 with b.
__device__ __forceinline__ int swzx(int p) { return p ^ (((p >> 6) & 3) << 1); }

template <bool SINGLE = false>
struct FftTx {  // M = 512, T = 64.  SINGLE: one shared-memory exchange buffer guarded by an mbarrier (see Fft<>)
  using G = Geo<9>;
  uint64_t* rd_bar = nullptr;
  uint32_t rd_phase = 0;
  double2 tl0, tl1, tl2, tl3;  // last-pass twiddles of block tau
  int parity;
  double2* ex;                 // [2][M] ping-pong buffers of the shared-memory exchange
  const Tw4* tab;
  int tau, lane, b, u;         // middle-pass role of this thread: block b, offset u
  uint32_t tmem;               // this warp's TMEM lane quadrant, 32 columns

  __device__ __forceinline__ void init(double2* ex_, const double2* tw_tab, int tau_, uint32_t tmem_) {
    ex = ex_; tab = reinterpret_cast<const Tw4*>(tw_tab); tau = tau_; lane = tau_ & 31; tmem = tmem_; parity = 0;
    b = ((tau_ >> 5) << 2) | ((lane >> 1) & 3);
    u = ((lane >> 3) << 1) | (lane & 1);
    const Tw4* e = tab + G::tab_off(2) + tau;
    tl0 = e->s[0]; tl1 = e->s[1]; tl2 = e->s[2]; tl3 = e->s[3];
  }
  __device__ __forceinline__ double2* next_buf() {
    if constexpr (SINGLE) {
      mbar_wait(rd_bar, rd_phase);  // every thread has finished reading the previous exchange
      rd_phase ^= 1u;
      return ex;
    } else {
      double2* buf = ex + (parity ? G::M : 0);
      parity ^= 1;
      return buf;
    }
  }
  __device__ __forceinline__ void done_reading() const {
    if constexpr (SINGLE) mbar_arrive(rd_bar);
  }

  // ---- exchange between pass 1 and pass 2 through TMEM + one lane-pair shuffle --------------------------------
  // in: x[a'] = point b*64 + u + 8a';  out: x[e] = point 8 tau + e
  __device__ __forceinline__ void exchange_tm_fwd(double2 (&x)[8]) const {
    uint32_t r[16];
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 8; a++) { r[2 * a] = (uint32_t)__double2loint(x[a].x); r[2 * a + 1] = (uint32_t)__double2hiint(x[a].x); }
    tmem_st16(tmem, r);
#pragma unroll
    for (int a = 0; a < 8; a++) { r[2 * a] = (uint32_t)__double2loint(x[a].y); r[2 * a + 1] = (uint32_t)__double2hiint(x[a].y); }
    tmem_st16(tmem + 16, r);
    tmem_wait_st();
    __syncwarp();
    uint32_t e[2][16];
    tmem_ld_16x256b_x4(tmem, e[0]);                 // source lanes t/4 (g = 0) and t/4 + 8 (g = 1)
    tmem_ld_16x256b_x4(tmem + (16u << 16), e[1]);   // source lanes t/4 + 16 (g = 2) and t/4 + 24 (g = 3)
    tmem_wait_ld();
    const bool t2 = (lane >> 2) & 1;
#pragma unroll
    for (int g = 0; g < 4; g++) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        // n = 2c + a'2 ; registers of (g, n): e[g >> 1][4n + 2 (g & 1) + {0, 1}]
        const int n0 = 2 * c, n1 = 2 * c + 1, o = 2 * (g & 1);
        const double E0 = __hiloint2double((int)e[g >> 1][4 * n0 + o + 1], (int)e[g >> 1][4 * n0 + o]);
        const double E1 = __hiloint2double((int)e[g >> 1][4 * n1 + o + 1], (int)e[g >> 1][4 * n1 + o]);
        const double send = t2 ? E0 : E1, keep = t2 ? E1 : E0;
        const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
        const double v0 = t2 ? recv : keep, v1 = t2 ? keep : recv;
        if (c == 0) { x[2 * g].x = v0; x[2 * g + 1].x = v1; } else { x[2 * g].y = v0; x[2 * g + 1].y = v1; }
      }
    }
  }
  // in: x[e] = point 8 tau + e;  out: x[a'] = point b*64 + u + 8a'
  __device__ __forceinline__ void exchange_tm_inv(double2 (&x)[8]) const {
    uint32_t e[2][16];
    const bool t2 = (lane >> 2) & 1;
#pragma unroll
    for (int g = 0; g < 4; g++) {
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const double a0 = c ? x[2 * g].y : x[2 * g].x, a1 = c ? x[2 * g + 1].y : x[2 * g + 1].x;
        const double keep = t2 ? a1 : a0, send = t2 ? a0 : a1;
        const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
        const double E0 = t2 ? recv : keep, E1 = t2 ? keep : recv;
        const int n0 = 2 * c, n1 = 2 * c + 1, o = 2 * (g & 1);
        e[g >> 1][4 * n0 + o] = (uint32_t)__double2loint(E0); e[g >> 1][4 * n0 + o + 1] = (uint32_t)__double2hiint(E0);
        e[g >> 1][4 * n1 + o] = (uint32_t)__double2loint(E1); e[g >> 1][4 * n1 + o + 1] = (uint32_t)__double2hiint(E1);
      }
    }
    __syncwarp();
    tmem_st_16x256b_x4(tmem, e[0]);
    tmem_st_16x256b_x4(tmem + (16u << 16), e[1]);
    tmem_wait_st();
    __syncwarp();
    uint32_t r[16], q[16];
    tmem_ld16(tmem, r);
    tmem_ld16(tmem + 16, q);
    tmem_wait_ld();
#pragma unroll
    for (int a = 0; a < 8; a++) {
      x[a].x = __hiloint2double((int)r[2 * a + 1], (int)r[2 * a]);
      x[a].y = __hiloint2double((int)q[2 * a + 1], (int)q[2 * a]);
    }
  }

  struct NoHook { __device__ __forceinline__ void operator()() const {} };
  __device__ __forceinline__ void forward(double2 (&x)[8], const Tw4& tw0) { forward(x, tw0, NoHook()); }
  // hook() runs right after the block barrier of the shared-memory exchange
  template <class Hook>
  __device__ __forceinline__ void forward(double2 (&x)[8], const Tw4& tw0, const Hook& hook) {
    radix8_fwd<3>(x, tw0.s[0], tw0.s[1], tw0.s[2], tw0.s[3]);
    {  // exchange 0 -> 1 through shared memory (crosses the two warps)
      double2* buf = next_buf();
#pragma unroll
      for (int a = 0; a < 8; a++) buf[swzx(tau + 64 * a)] = x[a];
      __syncthreads();
      hook();
      const int rb = 64 * b + u;
#pragma unroll
      for (int a = 0; a < 8; a++) x[a] = buf[swzx(rb + 8 * a)];
      done_reading();
    }
    {
      const Tw4* e = tab + G::tab_off(1) + b;
      const double2 s0 = __ldg(&e->s[0]), s1 = __ldg(&e->s[1]), s2 = __ldg(&e->s[2]), s3 = __ldg(&e->s[3]);
      radix8_fwd<3>(x, s0, s1, s2, s3);
    }
    exchange_tm_fwd(x);
    radix8_fwd<3>(x, tl0, tl1, tl2, tl3);
  }
  __device__ __forceinline__ void inverse(double2 (&x)[8], const Tw4& tw0) {
    radix8_inv<3>(x, tl0, tl1, tl2, tl3);
    exchange_tm_inv(x);
    {
      const Tw4* e = tab + G::tab_off(1) + b;
      const double2 s0 = __ldg(&e->s[0]), s1 = __ldg(&e->s[1]), s2 = __ldg(&e->s[2]), s3 = __ldg(&e->s[3]);
      radix8_inv<3>(x, s0, s1, s2, s3);
    }
    {
      double2* buf = next_buf();
      const int wb = 64 * b + u;
#pragma unroll
      for (int a = 0; a < 8; a++) buf[swzx(wb + 8 * a)] = x[a];
      __syncthreads();
#pragma unroll
      for (int a = 0; a < 8; a++) x[a] = buf[swzx(tau + 64 * a)];
      done_reading();
    }
    radix8_inv<3>(x, tw0.s[0], tw0.s[1], tw0.s[2], tw0.s[3]);
  }
};

template <int L, int BGBIT, bool SMALL, int MINB>
__global__ void __launch_bounds__(64, MINB) blind_rotate_tx_kernel(const BrArgs A) {
  constexpr int LOGN = 10, N = 1024, M = 512, T = 64;
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_tmem_base;
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                    // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);               // [2][M]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + 32 * M);
  const int tau = threadIdx.x;
  const long long g = blockIdx.x;
  const int n = A.n;
  const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);

  if (tau < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  FftTx<false> fft;
  fft.init(ex, A.tw_tab, tau, s_tmem_base + ((uint32_t)(tau & ~31) << 16));

  const size_t row_stride = (size_t)2 * L * 2 * M;
  for (int i = 0; i < n; i++) {
    const int at = abar[i];
    if (at == 0) continue;
    const double2* __restrict__ bk = A.bsk + row_stride * i + tau;
    double2 accA[8], accB[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { accA[e] = make_double2(0.0, 0.0); accB[e] = make_double2(0.0, 0.0); }
    TFHE_UNROLL(TFHE_TX_UNROLL_POLY)
    for (int poly = 0; poly < 2; poly++) {
      const uint32_t* P = acc + poly * N;
      uint32_t dre[8], dim[8];
      const int ib = (tau - at) & (2 * N - 1);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        dre[a] = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
        dim[a] = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
      }
      TFHE_UNROLL(TFHE_TX_UNROLL_LVL)
      for (int lvl = 0; lvl < L; lvl++) {
        double2 x[8];
        const int sh = 32 - (lvl + 1) * BGBIT;
#pragma unroll
        for (int a = 0; a < 8; a++) {
          x[a].x = digit_scaled<BGBIT>(dre[a], sh);
          x[a].y = digit_scaled<BGBIT>(dim[a], sh);
        }
        fft.forward(x, A.tw0);
        const double2* __restrict__ rowA = bk + (size_t)((poly * L + lvl) * 2) * M;
        const double2* __restrict__ rowB = rowA + M;
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const double2 ka = __ldg(rowA + e * T);
          const double2 kb = __ldg(rowB + e * T);
          accA[e].x = fma(x[e].x, ka.x, accA[e].x);
          accA[e].x = fma(-x[e].y, ka.y, accA[e].x);
          accA[e].y = fma(x[e].x, ka.y, accA[e].y);
          accA[e].y = fma(x[e].y, ka.x, accA[e].y);
          accB[e].x = fma(x[e].x, kb.x, accB[e].x);
          accB[e].x = fma(-x[e].y, kb.y, accB[e].x);
          accB[e].y = fma(x[e].x, kb.y, accB[e].y);
          accB[e].y = fma(x[e].y, kb.x, accB[e].y);
        }
      }
    }
    fft.inverse(accA, A.tw0);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      acc[j] += to_torus<SMALL>(accA[a].x);
      acc[j + M] += to_torus<SMALL>(accA[a].y);
    }
    fft.inverse(accB, A.tw0);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      acc[N + j] += to_torus<SMALL>(accB[a].x);
      acc[N + j + M] += to_torus<SMALL>(accB[a].y);
    }
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
  if (tau < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(s_tmem_base) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// "tmex" + TMA: with the second exchange off the LSU pipe there is room to stage the key rows through shared memory
// (cp.async.bulk + mbarrier, double-buffered, one transform ahead): the MAC then reads the key with short-latency
// LDS instead of waiting ~600 cycles on L2, and no registers are spent on keeping 16 loads in flight.
// ---------------------------------------------------------------------------------------------------------------
template <int L, int BGBIT, bool SMALL, int MINB>
__global__ void __launch_bounds__(64, MINB) blind_rotate_txs_kernel(const BrArgs A) {
  constexpr int LOGN = 10, N = 1024, M = 512, T = 64;
  constexpr uint32_t ROW_BYTES = 2u * M * 16u;
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_nsteps;
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                         // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);                    // [M]
  double2* kbuf = reinterpret_cast<double2*>(smem_raw + 8 * N + 16 * M);         // [2 buffers][2][8][T]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + 16 * M + 64 * M);
  const int tau = threadIdx.x;
  const long long g = blockIdx.x;
  const int n = A.n;
  unsigned short* steps = abar + n;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + 8 * N + 16 * M + 64 * M + (((n + 1) * 4 + 15) / 16 * 16));
  uint64_t* full = mbar;
  uint64_t* rd_bar = mbar + 2;
  const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);

  if (tau < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
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
  if (tau == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(rd_bar, T);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  FftTx<true> fft;
  fft.init(ex, A.tw_tab, tau, s_tmem_base + ((uint32_t)(tau & ~31) << 16));
  fft.rd_bar = rd_bar;
  mbar_arrive(rd_bar);  // phase 0: nothing outstanding before the first exchange
  if (tau == 0) {
    int c = 0;
    for (int i = 0; i < n; i++)
      if (abar[i] != 0) steps[c++] = (unsigned short)i;
    s_nsteps = c;
  }
  __syncthreads();
  const int nsteps = s_nsteps;
  const int njobs = nsteps * 2 * L;
  const size_t row_stride = (size_t)2 * L * 2 * M;
  const char* bsk_bytes = reinterpret_cast<const char*>(A.bsk);
  auto issue = [&](int q) {
    const int k = q / (2 * L), r = q - k * (2 * L);
    uint64_t* fb = &full[q & 1];
    mbar_arrive_expect_tx(fb, ROW_BYTES);
    bulk_copy_g2s(kbuf + (size_t)(q & 1) * 2 * M,
                  bsk_bytes + ((size_t)steps[k] * row_stride + (size_t)r * 2 * M) * sizeof(double2), ROW_BYTES, fb);
  };
  if (tau == 0 && njobs > 0) issue(0);
  int q = 0;

  for (int k = 0; k < nsteps; k++) {
    const int at = abar[steps[k]];
    double2 accA[8], accB[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { accA[e] = make_double2(0.0, 0.0); accB[e] = make_double2(0.0, 0.0); }
    TFHE_UNROLL(TFHE_TX_UNROLL_POLY)
    for (int poly = 0; poly < 2; poly++) {
      const uint32_t* P = acc + poly * N;
      uint32_t dre[8], dim[8];
      const int ib = (tau - at) & (2 * N - 1);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        dre[a] = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
        dim[a] = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
      }
      TFHE_UNROLL(TFHE_TX_UNROLL_LVL)
      for (int lvl = 0; lvl < L; lvl++, q++) {
        double2 x[8];
        const int sh = 32 - (lvl + 1) * BGBIT;
#pragma unroll
        for (int a = 0; a < 8; a++) {
          x[a].x = digit_scaled<BGBIT>(dre[a], sh);
          x[a].y = digit_scaled<BGBIT>(dim[a], sh);
        }
        fft.forward(x, A.tw0, [&]() { if (tau == 0 && q + 1 < njobs) issue(q + 1); });
        mbar_wait(&full[q & 1], (uint32_t)(q >> 1) & 1u);
        const double2* rowA = kbuf + (size_t)(q & 1) * 2 * M + tau;
        const double2* rowB = rowA + M;
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const double2 ka = rowA[e * T];
          const double2 kb = rowB[e * T];
          accA[e].x = fma(x[e].x, ka.x, accA[e].x);
          accA[e].x = fma(-x[e].y, ka.y, accA[e].x);
          accA[e].y = fma(x[e].x, ka.y, accA[e].y);
          accA[e].y = fma(x[e].y, ka.x, accA[e].y);
          accB[e].x = fma(x[e].x, kb.x, accB[e].x);
          accB[e].x = fma(-x[e].y, kb.y, accB[e].x);
          accB[e].y = fma(x[e].x, kb.y, accB[e].y);
          accB[e].y = fma(x[e].y, kb.x, accB[e].y);
        }
      }
    }
    fft.inverse(accA, A.tw0);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      acc[j] += to_torus<SMALL>(accA[a].x);
      acc[j + M] += to_torus<SMALL>(accA[a].y);
    }
    fft.inverse(accB, A.tw0);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      acc[N + j] += to_torus<SMALL>(accB[a].x);
      acc[N + j + M] += to_torus<SMALL>(accB[a].y);
    }
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
  if (tau < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(s_tmem_base) : "memory");
}

}  // namespace tfhe
