// blind_rotate_mg.cuh — several gates per thread block sharing ONE staged copy of the key rows (variant "mg").
//
// Same arithmetic and results as blind_rotate_kernel (blind_rotate.cuh; reference evaluator/evaluator.go:110-135).
// Motivation (profiles/r01_experiments.md): the block-per-gate kernel is limited by the number of resident warps
// (2 per scheduler at 255 registers).  Six gates per SM need <= 168 registers per thread, which rules out key rows in
// flight in registers; L1 prefetch cannot replace them (six independent gates stream 2 x 96 KiB per step through
// ~100 KiB of L1), and a private shared-memory stage per gate puts every key byte on the shared-memory pipe twice.
// Here G gates (G x N/16 threads) form one block and walk the bootstrapping key TOGETHER: the rows of every
// (step, digit) are copied once per block into a double-buffered shared-memory stage by a bulk asynchronous copy
// (cp.async.bulk + mbarrier, SASS UBLKCP) and read by all G gates, so the stage costs 1/G of a write per gate and the
// L2 -> SM traffic drops G-fold.  Nobody waits for the copy to be issued: each warp counts itself out of a stage
// after its multiply-accumulate, and the warp that happens to be last issues the copy of the digit two ahead.
// The spectrum accumulators live in TMEM (64 columns per gate, tcgen05.ld/st), each gate's transforms synchronise
// on its own named barrier, so the gates drift by up to one digit against each other (their shared-memory and FP64
// phases overlap) but never more.
// Shared memory: G x (8 KiB accumulator + 8 KiB exchange + 1.4 KiB) + 32 KiB stage; TMEM: 64 x ceil(G/2) columns.
#pragma once
#include "blind_rotate.cuh"
#include "blind_rotate_w16.cuh"  // TMEM primitives

namespace tfhe {

#ifndef TFHE_MG_SINGLE
#define TFHE_MG_SINGLE 0   // 1: one exchange buffer per gate guarded by an mbarrier; 0: two buffers used alternately (measured faster)
#endif
template <int LOGN, int G>
__host__ __device__ constexpr size_t br_mg_gate_bytes(int n) {
  return (size_t)8 * (1 << LOGN) /*acc*/ + (size_t)(TFHE_MG_SINGLE ? 1 : 2) * (1 << (LOGN - 1)) * 16 /*exchange*/ +
         (size_t)(((n + 1) * 2 + 15) / 16 * 16) /*abar*/;
}
template <int LOGN, int G>
constexpr size_t br_mg_smem_bytes(int n) {
  return (size_t)G * br_mg_gate_bytes<LOGN, G>(n) + (size_t)2 * 2 * (1 << (LOGN - 1)) * 16 /*2 key stages*/ + 128 /*mbarriers, counters*/;
}

template <int LOGN, int L, int BGBIT, bool SMALL, int G>
__global__ void __launch_bounds__(G * (1 << (LOGN - 4)), 6 / G) blind_rotate_mg_kernel(const BrArgs A, long long count) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  static_assert(T == 64 && (G == 2 || G == 3 || G == 6), "two warps per gate; 6 gates per SM");
  constexpr uint32_t ROW_BYTES = 2u * M * 16u;
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
  constexpr int TCOLS = (G <= 2) ? 64 : (G <= 4 ? 128 : 256);  // 64 columns per pair of gates (lane quadrants 0,1 / 2,3)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_tmem_base;
  const int n = A.n;
  const size_t gate_bytes = br_mg_gate_bytes<LOGN, G>(n);
  const int gate = threadIdx.x / T;  // gate of this block served by this thread
  const int tau = threadIdx.x % T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long g = (long long)blockIdx.x * G + gate;
  const int nact = (int)((count - (long long)blockIdx.x * G) < G ? (count - (long long)blockIdx.x * G) : G);  // gates of this block
  unsigned char* mine = smem_raw + gate * gate_bytes;
  uint32_t* acc = reinterpret_cast<uint32_t*>(mine);                                // [2][N]
  double2* ex = reinterpret_cast<double2*>(mine + 8 * N);                           // [M]
  unsigned short* abar = reinterpret_cast<unsigned short*>(mine + 8 * N + (TFHE_MG_SINGLE ? 16 : 32) * M);  // [n]
  unsigned char* shared_part = smem_raw + G * gate_bytes;
  double2* kbuf = reinterpret_cast<double2*>(shared_part);                          // [2 stages][2][8][T]
  uint64_t* full = reinterpret_cast<uint64_t*>(shared_part + 64 * M);               // [2]: stage b has landed
  uint64_t* rd_bar = full + 2;                                                      // [G]: exchange-buffer reads done
  uint64_t* empty = rd_bar + G;                                                     // [2]: every active warp is done with stage b
  int* cnt = reinterpret_cast<int*>(empty + 2);                                     // [2]: elects the warp that re-arms stage b
  const int njobs = n * 2 * L;  // job q = (step q / 2L, digit q % 2L): ALL steps, also those a gate skips
  const char* bsk_bytes = reinterpret_cast<const char*>(A.bsk);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    for (int j = 0; j < G; j++) mbar_init(&rd_bar[j], T);
    mbar_init(&empty[0], (uint32_t)(nact * (T / 32)));
    mbar_init(&empty[1], (uint32_t)(nact * (T / 32)));
    cnt[0] = 0;
    cnt[1] = 0;
  }
  const bool active = gate < nact;
  if (active) {
    const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);
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
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;
  if (threadIdx.x == 0) {  // the first two stages
    for (int q = 0; q < 2 && q < njobs; q++) {
      mbar_arrive_expect_tx(&full[q], ROW_BYTES);
      bulk_copy_g2s(kbuf + (size_t)q * 2 * M, bsk_bytes + (size_t)q * ROW_BYTES, ROW_BYTES, &full[q]);
    }
  }
  if (active) {
    Fft<LOGN - 1, TFHE_MG_SINGLE != 0, true> fft;
    fft.init(ex, A.tw_tab, tau);
    fft.bar_id = 1 + gate;
    if constexpr (TFHE_MG_SINGLE != 0) {
      fft.init_single(&rd_bar[gate]);
      mbar_arrive(&rd_bar[gate]);  // completes phase 0: "no reads outstanding" before the first exchange
    }
    // a warp may only touch TMEM lanes [32 (warp % 4), +32): gates 2j and 2j+1 use the four quadrants of columns [64 j, 64 j + 64)
    const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(64 * (gate >> 1));
    const int nwarps_active = nact * (T / 32);

    // done with stage (q & 1) of job q: the warp that is last re-arms it with job q + 2
    auto release_stage = [&](int q) {
      __syncwarp();
      if (lane == 0) {
        const int b = q & 1;
        mbar_arrive(&empty[b]);                 // release: this warp's reads of the stage are done
        const int old = atomicAdd(&cnt[b], 1);  // election only: exactly one warp sees the last ticket
        if (old == nwarps_active - 1) {
          cnt[b] = 0;
          if (q + 2 < njobs) {
            mbar_wait(&empty[b], (uint32_t)(q >> 1) & 1u);  // acquire: already complete (every arrive precedes its ticket)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the stage before the async-proxy write
            mbar_arrive_expect_tx(&full[b], ROW_BYTES);
            bulk_copy_g2s(kbuf + (size_t)b * 2 * M, bsk_bytes + (size_t)(q + 2) * ROW_BYTES, ROW_BYTES, &full[b]);
          }
        }
      }
    };

    int q = 0;
    for (int i = 0; i < n; i++) {
      const int at = abar[i];
      if (at == 0) {  // X^0: exact no-op for this gate; it still has to walk the shared stages in order
        for (int r = 0; r < 2 * L; r++, q++) {
          mbar_wait(&full[q & 1], (uint32_t)(q >> 1) & 1u);
          release_stage(q);
        }
        continue;
      }
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
          fft.forward(x, A.tw0);
          mbar_wait(&full[q & 1], (uint32_t)(q >> 1) & 1u);
          const double2* rowA = kbuf + (size_t)(q & 1) * 2 * M + tau;
          const double2* rowB = rowA + M;
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
              for (int c = 0; c < 4; c++) { aA[c] = make_double2(0.0, 0.0); aB[c] = make_double2(0.0, 0.0); }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
              const int e = 4 * h + c;
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
            pack4(aA, ra);
            pack4(aB, rb);
            tmem_st16(tacc + 16 * h, ra);
            tmem_st16(tacc + 32 + 16 * h, rb);
          }
          release_stage(q);
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
        fft.inverse(x, A.tw0);
        uint32_t* P = acc + poly * N;
#pragma unroll
        for (int a = 0; a < 8; a++) {
          const int j = tau + T * a;
          P[j] += to_torus<SMALL>(x[a].x);
          P[j + M] += to_torus<SMALL>(x[a].y);
        }
      }
      fft.group_barrier();  // accumulator updates visible to the gate's other warp before the next step reads them
    }

    if (A.out_mode == 0) {
      uint32_t* o = A.out + g * (2 * N);
      for (int j = tau; j < 2 * N; j += T) o[j] = acc[j];
    } else {  // sample extract at 0 (trlwe_ops.go:10-21)
      uint32_t* o = A.out + g * (N + 1);
      for (int j = tau; j < N; j += T) o[j] = (j == 0) ? acc[0] : ~acc[N - j];
      if (tau == 0) o[N] = acc[N];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS) : "memory");
}

}  // namespace tfhe
