// blind_rotate_tms.cuh — block-per-gate blind rotation built for SIX resident blocks per SM (variant "tms").
//
// Same arithmetic, mapping and results as blind_rotate_kernel (blind_rotate.cuh; reference evaluator/evaluator.go:110-135),
// but nothing long-lived sits in registers:
//   * the two spectrum accumulators (64 registers) live in 64 TMEM columns of the thread's lane and are
//     read-modify-written by the multiply-accumulate (tcgen05.ld / tcgen05.st, SASS LDTM / STTM);
//   * the key rows of the current digit are brought into ONE 2*M*16-byte shared-memory buffer by a bulk asynchronous copy
//     (cp.async.bulk + mbarrier, SASS UBLKCP) issued by one thread as soon as the previous digit's MAC is provably done
//     (right after the first block barrier of the next transform), so no registers are held in flight for the L2 latency
//     and nothing depends on L1 capacity (six blocks would stream 2 x 96 KiB per step through 88 KiB of L1);
//   * one exchange buffer, its write-after-read hazard covered by an mbarrier (Fft<.., SINGLE = true>).
// The register file is split per scheduler (16 K registers each), so the occupancy steps for 2-warp blocks are
// 4 blocks (255 registers), 6 blocks (168) and 8 blocks (128): this kernel targets 168.
// Shared memory per block at N = 1024: 8 KiB accumulator + 8 KiB exchange + 16 KiB key stage + ~1.5 KiB = 33.5 KiB.
#pragma once
#include "blind_rotate.cuh"
#include "blind_rotate_w16.cuh"  // TMEM primitives

namespace tfhe {

template <int LOGN>
constexpr size_t br_tms_smem_bytes(int n) {
  return (size_t)8 * (1 << LOGN) /*acc*/ + (size_t)(1 << (LOGN - 1)) * 16 /*exchange (single)*/ +
         (size_t)2 * (1 << (LOGN - 1)) * 16 /*key stage: A and B spectra of one digit*/ +
         (size_t)(((n + 1) * 4 + 15) / 16 * 16) /*abar + steps*/ + 32 /*mbarriers*/;
}

#ifndef TFHE_TMS_CHUNK
#define TFHE_TMS_CHUNK 4   // spectrum points per TMEM read-modify-write chunk (4 or 2)
#endif

template <int LOGN, int L, int BGBIT, bool SMALL, int MINB>
__global__ void __launch_bounds__((1 << (LOGN - 4)), MINB) blind_rotate_tms_kernel(const BrArgs A) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  static_assert(T >= 32 && T <= 128, "one TMEM lane per thread");
  constexpr uint32_t ROW_BYTES = 2u * M * 16u;
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_nsteps;
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                          // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);                     // [M]
  double2* kbuf = reinterpret_cast<double2*>(smem_raw + 8 * N + 16 * M);          // [2][8][T]: A then B spectrum rows
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + 16 * M + 32 * M);
  const int tau = threadIdx.x;
  const long long g = blockIdx.x;
  const int n = A.n;
  unsigned short* steps = abar + n;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + 8 * N + 16 * M + 32 * M + (((n + 1) * 4 + 15) / 16 * 16));
  uint64_t* full = mbar;        // the key stage has landed
  uint64_t* rd_bar = mbar + 1;  // exchange-buffer reads done
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
  if (tau == 0) {
    mbar_init(full, 1);
    mbar_init(rd_bar, T);
  }
  Fft<LOGN - 1, true> fft;
  fft.init(ex, A.tw_tab, tau);
  fft.init_single(rd_bar);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  mbar_arrive(rd_bar);  // completes phase 0: "no reads outstanding" before the first exchange
  if (tau == 0) {       // X^0 steps are exact no-ops (digits all zero): drop them from the schedule
    int c = 0;
    for (int i = 0; i < n; i++)
      if (abar[i] != 0) steps[c++] = (unsigned short)i;
    s_nsteps = c;
  }
  __syncthreads();
  const uint32_t tacc = s_tmem_base + ((uint32_t)(tau & ~31) << 16);  // columns [0,32) = A accumulator, [32,64) = B
  const int nsteps = s_nsteps;
  const int njobs = nsteps * 2 * L;  // job q = (schedule entry q / 2L, digit q % 2L)
  const size_t row_stride = (size_t)2 * L * 2 * M;
  const char* bsk_bytes = reinterpret_cast<const char*>(A.bsk);
  int issued = 0;  // jobs whose copy has been issued (meaningful in thread 0)
  // Called right after a block barrier: every thread has finished the MAC of job q - 1, so the stage is free for job q.
  auto stage_job = [&](int q) {
    if (tau == 0 && issued <= q && q < njobs) {
      const int k = q / (2 * L), r = q - k * (2 * L);
      mbar_arrive_expect_tx(full, ROW_BYTES);
      bulk_copy_g2s(kbuf, bsk_bytes + ((size_t)steps[k] * row_stride + (size_t)r * 2 * M) * sizeof(double2), ROW_BYTES, full);
      issued = q + 1;
    }
  };
  int q = 0;

  for (int k = 0; k < nsteps; k++) {
    const int at = abar[steps[k]];
#pragma unroll 1
    for (int poly = 0; poly < 2; poly++) {
      const uint32_t* P = acc + poly * N;
      uint32_t dre[8], dim[8];
      int ib = (tau - at) & (2 * N - 1);
      asm volatile("" : "+r"(ib));  // keep the rotated indices from being hoisted out of the loop and spilled
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        dre[a] = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
        dim[a] = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
      }
#pragma unroll 1
      for (int lvl = 0; lvl < L; lvl++, q++) {
        const int r = poly * L + lvl;
        const int sh = 32 - (lvl + 1) * BGBIT;
        double2 x[8];
#pragma unroll
        for (int a = 0; a < 8; a++) {
          x[a].x = digit_scaled<BGBIT>(dre[a], sh);
          x[a].y = digit_scaled<BGBIT>(dim[a], sh);
        }
        fft.forward(x, A.tw0, [&]() { stage_job(q); });
        mbar_wait(full, (uint32_t)q & 1u);
        const double2* rowA = kbuf + tau;
        const double2* rowB = rowA + M;
        constexpr int CH = TFHE_TMS_CHUNK;
#pragma unroll
        for (int h = 0; h < 8 / CH; h++) {
          double2 aA[CH], aB[CH];
          if constexpr (CH == 4) {
            uint32_t ra[16], rb[16];
            if (r > 0) {
              tmem_ld16(tacc + 16 * h, ra);
              tmem_ld16(tacc + 32 + 16 * h, rb);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < CH; c++) {
                aA[c] = make_double2(__hiloint2double((int)ra[4 * c + 1], (int)ra[4 * c]), __hiloint2double((int)ra[4 * c + 3], (int)ra[4 * c + 2]));
                aB[c] = make_double2(__hiloint2double((int)rb[4 * c + 1], (int)rb[4 * c]), __hiloint2double((int)rb[4 * c + 3], (int)rb[4 * c + 2]));
              }
            } else {
#pragma unroll
              for (int c = 0; c < CH; c++) { aA[c] = make_double2(0.0, 0.0); aB[c] = make_double2(0.0, 0.0); }
            }
          } else {
            uint32_t ra[8], rb[8];
            if (r > 0) {
              tmem_ld8(tacc + 8 * h, ra);
              tmem_ld8(tacc + 32 + 8 * h, rb);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 2; c++) {
                aA[c] = make_double2(__hiloint2double((int)ra[4 * c + 1], (int)ra[4 * c]), __hiloint2double((int)ra[4 * c + 3], (int)ra[4 * c + 2]));
                aB[c] = make_double2(__hiloint2double((int)rb[4 * c + 1], (int)rb[4 * c]), __hiloint2double((int)rb[4 * c + 3], (int)rb[4 * c + 2]));
              }
            } else {
#pragma unroll
              for (int c = 0; c < CH; c++) { aA[c] = make_double2(0.0, 0.0); aB[c] = make_double2(0.0, 0.0); }
            }
          }
#pragma unroll
          for (int c = 0; c < CH; c++) {
            const int e = CH * h + c;
            const double2 ka = rowA[e * T];
            const double2 kb = rowB[e * T];
            aA[c].x = fma(x[e].x, ka.x, aA[c].x);
            aA[c].x = fma(-x[e].y, ka.y, aA[c].x);
            aA[c].y = fma(x[e].x, ka.y, aA[c].y);
            aA[c].y = fma(x[e].y, ka.x, aA[c].y);
            aB[c].x = fma(x[e].x, kb.x, aB[c].x);
            aB[c].x = fma(-x[e].y, kb.y, aB[c].x);
            aB[c].y = fma(x[e].x, kb.y, aB[c].y);
            aB[c].y = fma(x[e].y, kb.x, aB[c].y);
          }
          if constexpr (CH == 4) {
            uint32_t ra[16], rb[16];
#pragma unroll
            for (int c = 0; c < CH; c++) {
              ra[4 * c] = (uint32_t)__double2loint(aA[c].x); ra[4 * c + 1] = (uint32_t)__double2hiint(aA[c].x);
              ra[4 * c + 2] = (uint32_t)__double2loint(aA[c].y); ra[4 * c + 3] = (uint32_t)__double2hiint(aA[c].y);
              rb[4 * c] = (uint32_t)__double2loint(aB[c].x); rb[4 * c + 1] = (uint32_t)__double2hiint(aB[c].x);
              rb[4 * c + 2] = (uint32_t)__double2loint(aB[c].y); rb[4 * c + 3] = (uint32_t)__double2hiint(aB[c].y);
            }
            tmem_st16(tacc + 16 * h, ra);
            tmem_st16(tacc + 32 + 16 * h, rb);
          } else {
            uint32_t ra[8], rb[8];
#pragma unroll
            for (int c = 0; c < 2; c++) {
              ra[4 * c] = (uint32_t)__double2loint(aA[c].x); ra[4 * c + 1] = (uint32_t)__double2hiint(aA[c].x);
              ra[4 * c + 2] = (uint32_t)__double2loint(aA[c].y); ra[4 * c + 3] = (uint32_t)__double2hiint(aA[c].y);
              rb[4 * c] = (uint32_t)__double2loint(aB[c].x); rb[4 * c + 1] = (uint32_t)__double2hiint(aB[c].x);
              rb[4 * c + 2] = (uint32_t)__double2loint(aB[c].y); rb[4 * c + 3] = (uint32_t)__double2hiint(aB[c].y);
            }
            tmem_st8(tacc + 8 * h, ra);
            tmem_st8(tacc + 32 + 8 * h, rb);
          }
        }
        tmem_wait_st();
      }
    }
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
        for (int c = 0; c < 4; c++) x[4 * h + c] = v[c];
      }
      // the first barrier inside the inverse transform proves the last MAC of this step is done everywhere:
      // stage the first digit of the NEXT step behind both inverse transforms
      fft.inverse(x, A.tw0, [&]() { stage_job(q); });
      uint32_t* P = acc + poly * N;
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        P[j] += to_torus<SMALL>(x[a].x);
        P[j + M] += to_torus<SMALL>(x[a].y);
      }
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
  if (tau < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(s_tmem_base) : "memory");
}

}  // namespace tfhe
