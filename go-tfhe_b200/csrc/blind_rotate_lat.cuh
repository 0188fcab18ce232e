// blind_rotate_lat.cuh — latency-oriented blind rotation for SMALL batches (variant "lat").
//
// The throughput kernel (blind_rotate.cuh) gives one gate two warps and relies on four co-resident gates per SM; a batch
// that does not even fill the SMs (a single gates.NAND call — BASELINE config 1 — or a narrow circuit level) leaves the
// machine idle while every gate walks its 6 forward + 2 inverse transforms per step one after the other.  Here a gate
// gets FOUR warps in two groups that work concurrently on the two polynomials of the accumulator:
//   group p (64 threads):  digits of polynomial p -> L forward transforms -> multiply-accumulate with rows p*L .. p*L+L-1
//                          into partial spectra for BOTH outputs                     (evaluator/evaluator.go:50-81)
//   exchange:              group 0 hands its partial B spectrum to group 1 and receives group 1's partial A spectrum
//   group p:               inverse transform of output p, rounding, accumulator update (poly/fourier_transform.go:88-125)
// so a step costs L forward + 1 inverse transform of latency instead of 2L + 2.  The partial sums are added in a
// different order than the reference's row-by-row accumulation, which is immaterial exactly where this kernel is
// enabled: the parameter sets whose rounded result is the exact integer result (SMALL: 80/110/128-bit; DESIGN.md
// section 2) — outputs are bit-identical to the default kernel and the oracle there.
#pragma once
#ifndef TFHE_BR_LAT_PF
#define TFHE_BR_LAT_PF 4   // two-group latency kernel: bulk L2 prefetch of the key row-set this many steps ahead (0 = off)
#endif
#include <cooperative_groups.h>

#include "blind_rotate.cuh"

namespace tfhe {

template <int LOGN>
constexpr size_t br_lat_smem_bytes(int n) {
  return (size_t)8 * (1 << LOGN) /*acc*/ + (size_t)2 * 2 * (1 << (LOGN - 1)) * 16 /*exchange: 2 groups x 2 buffers*/ +
         (size_t)2 * (1 << (LOGN - 1)) * 16 /*partial spectra crossing between the groups*/ +
         (size_t)2 * 2 * (1 << (LOGN - 1)) * 16 /*key rows of the next digit, one private slot set per thread*/ +
         (size_t)(((n + 1) * 2 + 15) / 16 * 16) /*abar*/;
}

template <int LOGN, int L, int BGBIT, bool SMALL>
__global__ void __launch_bounds__(2 * (1 << (LOGN - 4)), 1) blind_rotate_lat_kernel(const BrArgs A) {
  static_assert(SMALL, "partial sums are reordered: exact parameter sets only");
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                              // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);                         // [2 groups][2][M]
  double2* cross = reinterpret_cast<double2*>(smem_raw + 8 * N + 64 * M);             // [2][M]: [p] is read by group p
  double2* stage = reinterpret_cast<double2*>(smem_raw + 8 * N + 64 * M + 32 * M);    // [2 groups][16][T]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + 64 * M + 32 * M + 64 * M);
  const int grp = threadIdx.x / T, tau = threadIdx.x % T;
  const long long g = blockIdx.x;
  const int n = A.n;
  const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);
  // The 16 key values a thread multiplies with are copied asynchronously (cp.async, SASS LDGSTS) into 16 shared-memory
  // slots that only this thread reads, one digit ahead: with a single gate on the SM nothing else hides the L2 latency
  // of the key rows, and no registers are held in flight.  Thread-private slots need no barrier, only wait_group.
  double2* my_stage = stage + (size_t)grp * 16 * T + tau;
  auto stage_keys = [&](const double2* rowset, int r) {  // row r of the row-set (r >= 2L runs into the next step's row-set)
    const double2* base = rowset + (size_t)(r / (2 * L)) * (2 * L * 2 * M);
    r %= 2 * L;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_stage + e * T)), "l"(base + key_pos<T>(r, 0, e, tau)) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_stage + (8 + e) * T)), "l"(base + key_pos<T>(r, 1, e, tau)) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (int i = threadIdx.x; i < n; i += 2 * T) abar[i] = (unsigned short)br_modswitch<LOGN>(ct[i], A.ms_log2k);
  const int btil = br_btilde<LOGN>(ct[n], A.ms_log2k);
  const uint32_t* __restrict__ tv = br_testvec<N>(A, g);
  for (int j = threadIdx.x; j < N; j += 2 * T) {
    const int idx = (j - btil) & (2 * N - 1);
    const uint32_t va = tv[idx & (N - 1)], vb = tv[N + (idx & (N - 1))];
    acc[j] = (idx & N) ? ~va : va;
    acc[N + j] = (idx & N) ? ~vb : vb;
  }
  Fft<LOGN - 1, false, true> fft;
  fft.init(ex + (size_t)grp * 2 * M, A.tw_tab, tau);
  fft.bar_id = 1 + grp;
  __syncthreads();

  const size_t row_stride = (size_t)2 * L * 2 * M;
  uint32_t* P = acc + grp * N;  // the polynomial this group decomposes AND the one it updates
  int staged = -1;              // step whose first digit is already on its way into the stage
  for (int i = 0; i < n; i++) {
    const int at = abar[i];
    if (at == 0) continue;  // uniform over the block
    const double2* __restrict__ bk = A.bsk + row_stride * i;
#if TFHE_BR_LAT_PF
    // At most two gates per SM and all of them at the same step: nobody has warmed the key, every row-set would come
    // straight from HBM behind the cp.async that needs it.  One block per step (round robin) asks the bulk-copy unit to bring
    // the row-set of TFHE_BR_LAT_PF steps ahead into L2: single 128-bit gate 2.41 -> 2.30 ms.  (The order-preserving kernel
    // below does not do this: measured +3.7 % at N = 2048, flat at N = 1024, L = 1.)
    if (threadIdx.x == 0 && i + TFHE_BR_LAT_PF < n && (unsigned)i % gridDim.x == blockIdx.x)
      prefetch_l2_bulk(A.bsk + row_stride * (i + TFHE_BR_LAT_PF), (uint32_t)(row_stride * sizeof(double2)));
#endif
    if (staged != i) {  // first step, or the one after a skipped step: drop whatever is in flight, then stage the right rows
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      stage_keys(bk, grp * L);
    }
    double2 accA[8], accB[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { accA[e] = make_double2(0.0, 0.0); accB[e] = make_double2(0.0, 0.0); }
    uint32_t dre[8], dim[8];
    const int ib = (tau - at) & (2 * N - 1);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      dre[a] = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
      dim[a] = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
    }
#pragma unroll 1
    for (int lvl = 0; lvl < L; lvl++) {
      const int sh = 32 - (lvl + 1) * BGBIT;
      double2 x[8];
#pragma unroll
      for (int a = 0; a < 8; a++) {
        x[a].x = digit_scaled<BGBIT>(dre[a], sh);
        x[a].y = digit_scaled<BGBIT>(dim[a], sh);
      }
      fft.forward(x, A.tw0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      double2 ka[8], kb[8];
#pragma unroll
      for (int e = 0; e < 8; e++) { ka[e] = my_stage[e * T]; kb[e] = my_stage[(8 + e) * T]; }
      // next digit of this group: next level, or the first level of the next step (row-sets are contiguous; the key
      // buffer has slack past the last step) — its L2 latency hides behind the next transform(s)
      stage_keys(bk, (lvl + 1 < L) ? grp * L + lvl + 1 : 2 * L + grp * L);
#pragma unroll
      for (int e = 0; e < 8; e++) {
        accA[e].x = fma(x[e].x, ka[e].x, accA[e].x);
        accA[e].x = fma(-x[e].y, ka[e].y, accA[e].x);
        accA[e].y = fma(x[e].x, ka[e].y, accA[e].y);
        accA[e].y = fma(x[e].y, ka[e].x, accA[e].y);
        accB[e].x = fma(x[e].x, kb[e].x, accB[e].x);
        accB[e].x = fma(-x[e].y, kb[e].y, accB[e].x);
        accB[e].y = fma(x[e].x, kb[e].y, accB[e].y);
        accB[e].y = fma(x[e].y, kb[e].x, accB[e].y);
      }
    }
    // hand the partial spectrum of the OTHER group's output across, keep and complete our own
    if (grp == 0) {
#pragma unroll
      for (int e = 0; e < 8; e++) cross[M + e * T + tau] = accB[e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) cross[e * T + tau] = accA[e];
    }
    __syncthreads();
    double2 y[8];
    if (grp == 0) {
#pragma unroll
      for (int e = 0; e < 8; e++) { const double2 o = cross[e * T + tau]; y[e] = make_double2(accA[e].x + o.x, accA[e].y + o.y); }
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) { const double2 o = cross[M + e * T + tau]; y[e] = make_double2(accB[e].x + o.x, accB[e].y + o.y); }
    }
    fft.inverse(y, A.tw0);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      P[j] += to_torus<SMALL>(y[a].x);
      P[j + M] += to_torus<SMALL>(y[a].y);
    }
    __syncthreads();  // both polynomials updated (and `cross` free) before the next step reads them
    staged = i + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  br_write_output<N>(A, g, acc, acc + N, (int)threadIdx.x, 2 * T);
}

// ---------------------------------------------------------------------------------------------------------------
// Order-preserving two-group kernel for the parameter sets whose f64 sums are NOT exact (L <= 2: Uint1-5, the programmable
// bootstraps): the two groups still transform the two polynomials concurrently, but the multiply-accumulate chain keeps
// the reference's row order — group 0 accumulates rows 0..L-1 from zero, hands BOTH partial spectra to group 1, which
// (having kept its L digit spectra in registers) continues the very same FMA chain with rows L..2L-1 and returns the A
// spectrum; then each group inverse-transforms one output.  Every floating-point operation and its operands equal the
// default kernel's, so the result is bit-identical for every parameter set (variant "latp").
// Per step: max(L forward + L MAC, L forward) + L MAC + 1 inverse instead of 2L forward + 2L MAC + 2 inverse.
// ---------------------------------------------------------------------------------------------------------------
template <int LOGN, int L>
constexpr size_t br_latp_smem_bytes(int n) {
  return (size_t)8 * (1 << LOGN) /*acc*/ + (size_t)2 * 2 * (1 << (LOGN - 1)) * 16 /*exchange: 2 groups x 2 buffers*/ +
         (size_t)3 * (1 << (LOGN - 1)) * 16 /*partial A, partial B, final A crossing between the groups*/ +
         (size_t)2 * L * 2 * (1 << (LOGN - 1)) * 16 /*staged key rows: 2 groups x L digits x (A,B)*/ +
         (size_t)(((n + 1) * 2 + 15) / 16 * 16) /*abar*/;
}

template <int LOGN, int L, int BGBIT, bool SMALL>
__global__ void __launch_bounds__(2 * (1 << (LOGN - 4)), 1) blind_rotate_latp_kernel(const BrArgs A) {
  static_assert(L <= 2, "group 1 keeps its L digit spectra in registers");
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                                // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);                           // [2 groups][2][M]
  double2* cross = reinterpret_cast<double2*>(smem_raw + 8 * N + 64 * M);               // [3][M]: partial A, partial B, final A
  double2* stage = reinterpret_cast<double2*>(smem_raw + 8 * N + 64 * M + 48 * M);      // [2 groups][L][16][T]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + 64 * M + 48 * M + (size_t)L * 64 * M);
  const int grp = threadIdx.x / T, tau = threadIdx.x % T;
  const long long g = blockIdx.x;
  const int n = A.n;
  const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);
  double2* my_stage = stage + (size_t)grp * L * 16 * T + tau;

  for (int i = threadIdx.x; i < n; i += 2 * T) abar[i] = (unsigned short)br_modswitch<LOGN>(ct[i], A.ms_log2k);
  const int btil = br_btilde<LOGN>(ct[n], A.ms_log2k);
  const uint32_t* __restrict__ tv = br_testvec<N>(A, g);
  for (int j = threadIdx.x; j < N; j += 2 * T) {
    const int idx = (j - btil) & (2 * N - 1);
    const uint32_t va = tv[idx & (N - 1)], vb = tv[N + (idx & (N - 1))];
    acc[j] = (idx & N) ? ~va : va;
    acc[N + j] = (idx & N) ? ~vb : vb;
  }
  Fft<LOGN - 1, false, true> fft;
  fft.init(ex + (size_t)grp * 2 * M, A.tw_tab, tau);
  fft.bar_id = 1 + grp;
  __syncthreads();

  const size_t row_stride = (size_t)2 * L * 2 * M;
  uint32_t* P = acc + grp * N;
  auto mac = [&](const double2 (&x)[8], int lvl, double2 (&accA)[8], double2 (&accB)[8]) {
    const double2* ks = my_stage + (size_t)lvl * 16 * T;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const double2 ka = ks[e * T], kb = ks[(8 + e) * T];
      accA[e].x = fma(x[e].x, ka.x, accA[e].x);
      accA[e].x = fma(-x[e].y, ka.y, accA[e].x);
      accA[e].y = fma(x[e].x, ka.y, accA[e].y);
      accA[e].y = fma(x[e].y, ka.x, accA[e].y);
      accB[e].x = fma(x[e].x, kb.x, accB[e].x);
      accB[e].x = fma(-x[e].y, kb.y, accB[e].x);
      accB[e].y = fma(x[e].x, kb.y, accB[e].y);
      accB[e].y = fma(x[e].y, kb.x, accB[e].y);
    }
  };
  for (int i = 0; i < n; i++) {
    const int at = abar[i];
    if (at == 0) continue;  // uniform over the block
    {  // this group's L x 16 key values of this step, asynchronously into thread-private slots (hidden behind the transforms)
      const double2* __restrict__ rows = A.bsk + row_stride * i;
#pragma unroll
      for (int l = 0; l < L; l++)
#pragma unroll
        for (int e = 0; e < 8; e++) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_stage + (size_t)(l * 16 + e) * T)),
                       "l"(rows + key_pos<T>(grp * L + l, 0, e, tau)) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_stage + (size_t)(l * 16 + 8 + e) * T)),
                       "l"(rows + key_pos<T>(grp * L + l, 1, e, tau)) : "memory");
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t dre[8], dim[8];
    const int ib = (tau - at) & (2 * N - 1);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      dre[a] = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
      dim[a] = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
    }
    double2 xs[L][8];
    double2 accA[8], accB[8];
#pragma unroll
    for (int lvl = 0; lvl < L; lvl++) {
      const int sh = 32 - (lvl + 1) * BGBIT;
#pragma unroll
      for (int a = 0; a < 8; a++) {
        xs[lvl][a].x = digit_scaled<BGBIT>(dre[a], sh);
        xs[lvl][a].y = digit_scaled<BGBIT>(dim[a], sh);
      }
      fft.forward(xs[lvl], A.tw0);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    double2 y[8];
    if (grp == 0) {
#pragma unroll
      for (int e = 0; e < 8; e++) { accA[e] = make_double2(0.0, 0.0); accB[e] = make_double2(0.0, 0.0); }
#pragma unroll
      for (int lvl = 0; lvl < L; lvl++) mac(xs[lvl], lvl, accA, accB);
#pragma unroll
      for (int e = 0; e < 8; e++) { cross[e * T + tau] = accA[e]; cross[M + e * T + tau] = accB[e]; }
    }
    __syncthreads();  // #1: partials of rows 0..L-1 are out
    if (grp == 1) {
#pragma unroll
      for (int e = 0; e < 8; e++) { accA[e] = cross[e * T + tau]; accB[e] = cross[M + e * T + tau]; }
#pragma unroll
      for (int lvl = 0; lvl < L; lvl++) mac(xs[lvl], lvl, accA, accB);
#pragma unroll
      for (int e = 0; e < 8; e++) { cross[2 * M + e * T + tau] = accA[e]; y[e] = accB[e]; }
    }
    __syncthreads();  // #2: group 1 has finished the chain
    if (grp == 0) {
#pragma unroll
      for (int e = 0; e < 8; e++) y[e] = cross[2 * M + e * T + tau];
    }
    fft.inverse(y, A.tw0);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      P[j] += to_torus<SMALL>(y[a].x);
      P[j + M] += to_torus<SMALL>(y[a].y);
    }
    __syncthreads();  // both polynomials updated, `cross` and the stage free
  }

  br_write_output<N>(A, g, acc, acc + N, (int)threadIdx.x, 2 * T);
}

// ---------------------------------------------------------------------------------------------------------------
// One group of N/16 threads per DIGIT (2L groups, 384 threads at L = 3): all 2L forward transforms of a CMUX step run
// concurrently, every group writes its two partial products (digit spectrum x key row, for the A and the B output) to
// shared memory, and the first group of each polynomial sums the 2L partials of "its" output in the reference's row order,
// inverse-transforms and updates the accumulator.  One forward + one inverse transform of latency per step; the key rows
// of the NEXT step are requested before the block barrier so that their L2 latency hides behind the inverse transform.
// For batches of at most one gate per SM (variant "lat2"; same exactness condition as above).
// ---------------------------------------------------------------------------------------------------------------
template <int LOGN, int L>
constexpr size_t br_lat2_smem_bytes(int n) {
  // the partial products reuse the groups' exchange buffers (free once the forward transform is done), which keeps
  // shared memory at ~105 KiB and leaves L1 room for the prefetched key rows of the next step (96 KiB)
  return (size_t)8 * (1 << LOGN) /*acc*/ + (size_t)2 * L * 2 * (1 << (LOGN - 1)) * 16 /*exchange / partials: 2L groups x 2 buffers*/ +
         (size_t)(((n + 1) * 2 + 15) / 16 * 16) /*abar*/;
}

template <int LOGN, int L, int BGBIT, bool SMALL>
__global__ void __launch_bounds__(2 * L * (1 << (LOGN - 4)), 1) blind_rotate_lat2_kernel(const BrArgs A) {
  static_assert(SMALL, "partial sums are reordered: exact parameter sets only");
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8, NG = 2 * L;
  static_assert(NG <= 15, "one named barrier per group");
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                                   // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);                              // [NG][2][M]
  double2* part = ex;                                                                      // [NG][2 outputs][M], aliases ex
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + (size_t)NG * 32 * M);
  const int grp = threadIdx.x / T, tau = threadIdx.x % T;
  const int poly = grp / L, lvl = grp % L;
  const long long g = blockIdx.x;
  const int n = A.n;
  const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);

  for (int i = threadIdx.x; i < n; i += NG * T) abar[i] = (unsigned short)((ct[i] + (1u << (30 - LOGN))) >> (31 - LOGN));
  const unsigned long long bb = (unsigned long long)ct[n] + (1ull << (30 - LOGN));
  const int btil = (int)((2 * N - (int)(bb >> (31 - LOGN))) & (2 * N - 1));
  const uint32_t* __restrict__ tv = A.luts ? A.luts + (A.nluts == 1 ? 0 : g) * (2 * N) : A.testvec;
  for (int j = threadIdx.x; j < N; j += NG * T) {
    const int idx = (j - btil) & (2 * N - 1);
    const uint32_t va = tv[idx & (N - 1)], vb = tv[N + (idx & (N - 1))];
    acc[j] = (idx & N) ? ~va : va;
    acc[N + j] = (idx & N) ? ~vb : vb;
  }
  Fft<LOGN - 1, false, true> fft;
  fft.init(ex + (size_t)grp * 2 * M, A.tw_tab, tau);
  fft.bar_id = 1 + grp;
  __syncthreads();

  const size_t row_stride = (size_t)2 * L * 2 * M;
  const int sh = 32 - (lvl + 1) * BGBIT;
  uint32_t* P = acc + poly * N;  // the polynomial this group takes its digit from (and, for lvl == 0, the one it updates)
  auto next_step = [&](int i) { while (i < n && abar[i] == 0) i++; return i; };  // X^0 steps are exact no-ops
  auto prefetch_keys = [&](int i) {  // this group's 16 KiB of step i into L1: 128 lines, two per thread
    const char* pf = reinterpret_cast<const char*>(A.bsk + row_stride * i + (size_t)(grp * 2) * M) + tau * 128;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + T * 128));
  };
  auto load_keys = [&](int i, double2 (&ka)[8], double2 (&kb)[8]) {
    const double2* __restrict__ rowA = A.bsk + row_stride * i + (size_t)(grp * 2) * M + tau;
    const double2* __restrict__ rowB = rowA + M;
#pragma unroll
    for (int e = 0; e < 8; e++) { ka[e] = __ldg(rowA + e * T); kb[e] = __ldg(rowB + e * T); }
  };
  double2 ka[8], kb[8];
  int i = next_step(0);
  if (i < n) load_keys(i, ka, kb);
  while (i < n) {
    const int at = abar[i];
    const int inext = next_step(i + 1);
    if (inext < n) prefetch_keys(inext);  // a whole step ahead of the loads that will want them
    double2 x[8];
    {
      const int ib = (tau - at) & (2 * N - 1);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        const uint32_t wre = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
        const uint32_t wim = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
        x[a].x = digit_scaled<BGBIT>(wre, sh);
        x[a].y = digit_scaled<BGBIT>(wim, sh);
      }
    }
    fft.forward(x, A.tw0);
    fft.group_barrier();  // every thread of the group has read the last exchange: both buffers are free for the partials
    double2* pa = part + (size_t)(grp * 2) * M + tau;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      pa[e * T] = make_double2(fma(x[e].x, ka[e].x, -(x[e].y * ka[e].y)), fma(x[e].x, ka[e].y, x[e].y * ka[e].x));
      pa[M + e * T] = make_double2(fma(x[e].x, kb[e].x, -(x[e].y * kb[e].y)), fma(x[e].x, kb[e].y, x[e].y * kb[e].x));
    }
    if (inext < n) load_keys(inext, ka, kb);  // L1 hits by now; in flight across the barrier and the inverse transform
    __syncthreads();
    if (lvl == 0) {  // groups 0 and L: output `poly` = sum over all 2L digits, in row order
      double2 y[8];
#pragma unroll
      for (int e = 0; e < 8; e++) y[e] = part[(size_t)poly * M + e * T + tau];
#pragma unroll
      for (int r = 1; r < NG; r++) {
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const double2 o = part[(size_t)(r * 2 + poly) * M + e * T + tau];
          y[e].x += o.x;
          y[e].y += o.y;
        }
      }
      // the inverse transform reuses this group's exchange buffers, which hold partials the OTHER summing group reads too
      asm volatile("bar.sync %0, %1;" ::"n"(NG + 1), "n"(2 * T) : "memory");
      fft.inverse(y, A.tw0);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        P[j] += to_torus<SMALL>(y[a].x);
        P[j + M] += to_torus<SMALL>(y[a].y);
      }
    }
    __syncthreads();  // both polynomials updated and `part` free before the next step
    i = inext;
  }

  if (A.out_mode == 0) {
    uint32_t* o = A.out + g * (2 * N);
    for (int j = threadIdx.x; j < 2 * N; j += NG * T) o[j] = acc[j];
  } else {
    uint32_t* o = A.out + g * (N + 1);
    for (int j = threadIdx.x; j < N; j += NG * T) o[j] = (j == 0) ? acc[0] : ~acc[N - j];
    if (threadIdx.x == 0) o[N] = acc[N];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One gate on a CLUSTER of 2L thread blocks (2L SMs), one block of N/16 threads per digit (variant "cl").
// The per-digit kernel above shares one SM's FP64 pipe and key bandwidth among its 2L transforms; here every digit has
// an SM of its own and the blocks talk through distributed shared memory (cooperative_groups::this_cluster):
//   block r:            digit r of polynomial r / L (from a local copy of that polynomial) -> forward transform ->
//                       partial products with key row r for both outputs, left in LOCAL shared memory; the key rows of
//                       the next step are staged by cp.async meanwhile
//   cluster barrier
//   blocks 0 and L:     read the 2L partials of output A resp. B out of the other blocks' shared memory, sum them in row
//                       order, inverse-transform, round, and write the updated polynomial into the local copies of the L
//                       blocks that decompose it
//   cluster barrier
// One forward + one inverse transform, two cluster barriers and ~100 KiB of DSMEM traffic per step.  For batches of at
// most (SMs / 2L) gates — the single gates.NAND call; exact parameter sets only (partial sums, not the FMA chain).
// ---------------------------------------------------------------------------------------------------------------
template <int LOGN>
constexpr size_t br_cl_smem_bytes(int n) {
  return (size_t)4 * (1 << LOGN) /*one polynomial of the accumulator*/ + (size_t)2 * (1 << (LOGN - 1)) * 16 /*exchange*/ +
         (size_t)2 * (1 << (LOGN - 1)) * 16 /*partial products for A and B*/ + (size_t)2 * (1 << (LOGN - 1)) * 16 /*key stage*/ +
         (size_t)(((n + 1) * 2 + 15) / 16 * 16) /*abar*/;
}

template <int LOGN, int L, int BGBIT, bool SMALL>
__global__ void __cluster_dims__(2 * L, 1, 1) __launch_bounds__((1 << (LOGN - 4)), 1) blind_rotate_cl_kernel(const BrArgs A) {
  static_assert(SMALL, "partial sums are reordered: exact parameter sets only");
  namespace cg = cooperative_groups;
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8, NG = 2 * L;
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* P = reinterpret_cast<uint32_t*>(smem_raw);                                   // [N]: polynomial rank / L
  double2* ex = reinterpret_cast<double2*>(smem_raw + 4 * N);                            // [2][M]
  double2* part = reinterpret_cast<double2*>(smem_raw + 4 * N + 32 * M);                 // [2 outputs][M]
  double2* stage = reinterpret_cast<double2*>(smem_raw + 4 * N + 64 * M);                // [16][T]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 4 * N + 96 * M);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int poly = rank / L, lvl = rank % L;
  const int tau = threadIdx.x;
  const long long g = blockIdx.x / NG;
  const int n = A.n;
  const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);

  for (int i = tau; i < n; i += T) abar[i] = (unsigned short)((ct[i] + (1u << (30 - LOGN))) >> (31 - LOGN));
  const unsigned long long bb = (unsigned long long)ct[n] + (1ull << (30 - LOGN));
  const int btil = (int)((2 * N - (int)(bb >> (31 - LOGN))) & (2 * N - 1));
  const uint32_t* __restrict__ tv = (A.luts ? A.luts + (A.nluts == 1 ? 0 : g) * (2 * N) : A.testvec) + poly * N;
  for (int j = tau; j < N; j += T) {
    const int idx = (j - btil) & (2 * N - 1);
    const uint32_t v = tv[idx & (N - 1)];
    P[j] = (idx & N) ? ~v : v;
  }
  Fft<LOGN - 1, false, false> fft;
  fft.init(ex, A.tw_tab, tau);
  double2* my_stage = stage + tau;
  const size_t row_stride = (size_t)2 * L * 2 * M;
  auto stage_keys = [&](int i) {
    const double2* __restrict__ rowA = A.bsk + row_stride * i + (size_t)(rank * 2) * M + tau;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_stage + e * T)), "l"(rowA + e * T) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_stage + (8 + e) * T)), "l"(rowA + M + e * T) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // remote views: the partial buffers of every block (for the two summing blocks) and the polynomial copies this block feeds
  double2* rpart[NG];
#pragma unroll
  for (int r = 0; r < NG; r++) rpart[r] = cluster.map_shared_rank(part, r);
  uint32_t* rP[L];
#pragma unroll
  for (int l = 0; l < L; l++) rP[l] = cluster.map_shared_rank(P, poly * L + l);
  __syncthreads();
  auto next_step = [&](int i) { while (i < n && abar[i] == 0) i++; return i; };  // X^0 steps are exact no-ops
  int i = next_step(0);
  if (i < n) stage_keys(i);
  cluster.sync();
  const int sh = 32 - (lvl + 1) * BGBIT;
  while (i < n) {
    const int at = abar[i];
    double2 x[8];
    {
      const int ib = (tau - at) & (2 * N - 1);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        const uint32_t wre = rot_read<N>(P, ib + T * a) - P[j] + A.offset;
        const uint32_t wim = rot_read<N>(P, ib + T * a + M) - P[j + M] + A.offset;
        x[a].x = digit_scaled<BGBIT>(wre, sh);
        x[a].y = digit_scaled<BGBIT>(wim, sh);
      }
    }
    fft.forward(x, A.tw0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const double2 ka = my_stage[e * T], kb = my_stage[(8 + e) * T];
      part[e * T + tau] = make_double2(fma(x[e].x, ka.x, -(x[e].y * ka.y)), fma(x[e].x, ka.y, x[e].y * ka.x));
      part[M + e * T + tau] = make_double2(fma(x[e].x, kb.x, -(x[e].y * kb.y)), fma(x[e].x, kb.y, x[e].y * kb.x));
    }
    const int inext = next_step(i + 1);
    if (inext < n) stage_keys(inext);  // hidden behind the barriers and the inverse transform
    cluster.sync();                    // every block's partial products are in its shared memory
    if (lvl == 0) {                    // blocks 0 and L: output `poly` = sum over all 2L digits, in row order
      double2 y[8];
#pragma unroll
      for (int e = 0; e < 8; e++) y[e] = rpart[0][(size_t)poly * M + e * T + tau];
#pragma unroll
      for (int r = 1; r < NG; r++) {
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const double2 o = rpart[r][(size_t)poly * M + e * T + tau];
          y[e].x += o.x;
          y[e].y += o.y;
        }
      }
      fft.inverse(y, A.tw0);
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        const uint32_t v0 = P[j] + to_torus<SMALL>(y[a].x), v1 = P[j + M] + to_torus<SMALL>(y[a].y);
#pragma unroll
        for (int l = 0; l < L; l++) { rP[l][j] = v0; rP[l][j + M] = v1; }  // l = 0 is this block's own copy
      }
    }
    cluster.sync();  // updated polynomials visible everywhere; partial buffers free
    i = inext;
  }
  if (lvl == 0) {
    if (A.out_mode == 0) {
      uint32_t* o = A.out + g * (2 * N) + poly * N;
      for (int j = tau; j < N; j += T) o[j] = P[j];
    } else {  // sample extract at 0: block 0 holds polynomial A, block L holds B[0]
      uint32_t* o = A.out + g * (N + 1);
      if (poly == 0) { for (int j = tau; j < N; j += T) o[j] = (j == 0) ? P[0] : ~P[N - j]; }
      else if (tau == 0) o[N] = P[0];
    }
  }
  cluster.sync();  // no block may exit while others can still address its shared memory
}

}  // namespace tfhe
