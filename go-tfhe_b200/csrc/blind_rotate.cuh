// blind_rotate.cuh — fused blind-rotation kernel (one thread block per gate) for sm_100a.
//
// Replaces, for a whole batch of independent ciphertexts, the reference's
//   Evaluator.BlindRotateAssign   evaluator/evaluator.go:110-135   (n CMUX steps)
//   Evaluator.CMuxAssign          evaluator/evaluator.go:85-106
//   Evaluator.ExternalProductAssign evaluator/evaluator.go:50-81
//   poly.DecomposePolyAssign      poly/decomposer.go:55-66
//   Evaluator.ToFourierPolyAssign / ToPolyAssignUnsafe / MulAddFourierPolyAssign
//                                 poly/fourier_transform.go:18-125,170-347, poly/fourier_ops.go:167-191
//   poly.PolyMulWithXKInPlace     poly/buffer_methods.go:133-164
//   trlwe.SampleExtractIndexAssign(k = 0)  trlwe/trlwe_ops.go:10-21  (epilogue)
//
// Design: the accumulator TRLWE lives in shared memory for all n steps; each of
// the T = N/16 threads owns 8 complex points of every transform in registers.  The negacyclic
// transform is the same factorisation the reference uses (split x^M - c into x^(M/2) -+ sqrt(c),
// so the fold twist is inside the twiddles and spectra come out in the reference's own order,
// which lets the bootstrapping key be used as uploaded), but run as radix-8 register passes with
// FMA butterflies and bank-conflict-free swizzled shared-memory exchanges between passes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

// Build-time tuning knobs (defaults are the measured best; see profiles/).
#ifndef TFHE_BR_UNROLL_POLY
#define TFHE_BR_UNROLL_POLY 2   // 2 = fully unrolled over the A/B polynomials, 1 = rolled
#endif
#ifndef TFHE_BR_UNROLL_LVL
#define TFHE_BR_UNROLL_LVL 1    // 1 = level loop rolled (measured best: smaller code, no spills), >= L = fully unrolled
#endif
#ifndef TFHE_BR_KEEP_OWN
#define TFHE_BR_KEEP_OWN 0      // exchanges keep the one point that does not change owner in its register: 1 = in every exchange between full passes (round 1: 3.5 % slower), 2 = only where the kept slot is the same for 8 consecutive lanes, i.e. where a whole quarter-warp wavefront disappears
#endif
#ifndef TFHE_BR_TL_RELOAD
#define TFHE_BR_TL_RELOAD 0     // 1: last-pass twiddles re-read from L1 every transform (frees 16 registers)
#endif
#ifndef TFHE_BR_SHFL_LAST
#define TFHE_BR_SHFL_LAST 1     // N = 2048: last radix-2 stage through a lane-pair shuffle instead of a third exchange
#endif
#ifndef TFHE_BR_PAIR_INV
#define TFHE_BR_PAIR_INV 0      // 1: the two inverse transforms of a step run interleaved, sharing their exchanges (measured: no gain)
#endif
#ifndef TFHE_BR_PAIR_FWD
#define TFHE_BR_PAIR_FWD 0      // forward transforms of consecutive decomposition levels run interleaved in pairs
#endif
#ifndef TFHE_BR_PF_L1
#define TFHE_BR_PF_L1 0         // d > 0: prefetch.global.L1 of the key rows d digits ahead of the MAC that uses them (2 lines per thread).  +0.5 % with the round-1 kernel, -3.7 % with the persistent work-item kernel (97.1 k -> 100.7 k gates/s without it): off
#endif
#ifndef TFHE_BR_TW_CONST
#define TFHE_BR_TW_CONST 0      // 1: twiddles of the passes with <= 8 blocks (pass 1) come from the constant bank (LDC, not an LSU instruction) instead of L1
#endif
#ifndef TFHE_BR_KO
#define TFHE_BR_KO 0            // TIMING EXPERIMENTS ONLY (wrong results): bit 0 = no exchanges, bit 1 = no key loads, bit 2 = no barrier in exchanges
#endif
// exchange buffers hold one transform ([2][M]) unless a paired mode needs two ([2][2M])
#ifndef TFHE_TM_PAIR
#define TFHE_TM_PAIR 0          // TMEM-accumulator kernel: forward transforms of two levels and the two inverse transforms run paired
#endif
#define TFHE_BR_EXW ((TFHE_BR_PAIR_INV || TFHE_BR_PAIR_FWD || TFHE_TM_PAIR) ? 2 : 1)
#ifndef TFHE_BR_WARP_EX
#define TFHE_BR_WARP_EX 0       // 1: exchanges whose 8-thread groups lie inside one warp use a third buffer and __syncwarp instead of a block barrier
#endif
#ifndef TFHE_BR_WARP_EX_LOGN
#define TFHE_BR_WARP_EX_LOGN 11 // ... and always from this ring size on (4-warp blocks: measured +2.6 % at N = 2048, -9 % at N = 1024)
#endif
__host__ __device__ constexpr bool br_warp_ex(int logn) { return TFHE_BR_WARP_EX || logn >= TFHE_BR_WARP_EX_LOGN; }
__host__ __device__ constexpr int br_nbuf(int logn) { return br_warp_ex(logn) ? 3 : 2; }  // exchange buffers per transform group
#ifndef TFHE_EXPERIMENTAL
#define TFHE_EXPERIMENTAL 0
#endif
#ifndef TFHE_BR_MINB_N1024
#define TFHE_BR_MINB_N1024 4    // resident blocks per SM the N = 1024, L = 3 throughput kernel is compiled for
#endif
// The shipped instantiations of the throughput kernel (one per parameter-set shape, params/params.go:83-391).  They are
// compiled in their own translation unit (blind_rotate_throughput.cu) because ptxas's --register-usage-level=7 is worth
// +0.7 % at 128-bit and +4.6 % at Uint3 for THIS kernel while it slows the latency kernels and the tiled key switch
// (profiles/r02_experiments.md); the engine's translation unit only declares them.
namespace tfhe { struct BrArgs; void (*blind_rotate_throughput_kernel(int logN, int L, int bgbit))(const BrArgs); }
#define TFHE_BR_THROUGHPUT_INSTANCES(X) \
  X(10, 3, 6, true, TFHE_BR_MINB_N1024)  \
  X(10, 2, 10, false, 4)                 \
  X(9, 1, 18, false, 8)                  \
  X(10, 1, 23, false, 4)                 \
  X(11, 1, 22, false, 2)
#define TFHE_PRAGMA_(x) _Pragma(#x)
#define TFHE_UNROLL(n) TFHE_PRAGMA_(unroll n)

namespace tfhe {

struct Tw4 { double2 s[4]; };  // twiddles of one radix-8 block: S(m,i), S(2m,2i), S(4m,4i), S(4m,4i+2)
#if TFHE_BR_TW_CONST
__constant__ Tw4 c_tw_pass1[3][8];  // [LOGM - 8][block] : pass-1 twiddles of the M = 256, 512, 1024 transforms
#endif

constexpr int BR_MAX_TAIL = 7;
struct BrArgs {
  const uint32_t* ct_in;    // [count][n+1]  (already linearly combined)
  const uint32_t* testvec;  // [2][N] default test vector
  const uint32_t* luts;     // NULL or [nluts][2][N]
  long long nluts;
  const double2* bsk;       // [n][2L][2][8][T], pre-scaled by 1/M
  cudaTextureObject_t bsk_tex;  // the same buffer as a linear uint4 texture (TEX-path variant)
  const double2* tw_tab;    // per-pass twiddle tables for passes >= 1 (4 double2 per block)
  uint32_t* out;            // out_mode 0: TRLWE [count][2][N]; 1: extracted LWE [count][N+1]
  int n;
  uint32_t offset;          // CloudKey.DecompositionOffset
  int out_mode;
  Tw4 tw0;                  // pass-0 twiddles (same for every thread; lives in the constant bank)
  const int* lut_index;     // NULL, or [count] indices into luts (a small LUT table shared by many ciphertexts)
  int ms_log2k;             // many-LUT bootstraps: the mod switch keeps multiples of 2^ms_log2k only (0 = reference)
  int extract_k;            // out_mode 2: samples extracted at indices 0..extract_k-1 -> out [count][extract_k][N+1]
  // work distribution of the persistent throughput kernel (blind_rotate_kernel); ignored by the latency kernels
  long long count;          // gates in this launch
  int nchunks;              // work items per gate (1 = a whole gate per item)
  int chunk_steps;          // CMUX steps per work item of the first main_chunks items
  int main_chunks;          // items of chunk_steps steps; the remaining nchunks - main_chunks items are the graded tail:
  int tail_start[BR_MAX_TAIL + 1];  // item main_chunks + k covers steps [tail_start[k], tail_start[k + 1]) — ever shorter
                            // items at the end of every gate, so that the blocks of a launch finish close together
  uint32_t* scratch;        // [count][2][N] accumulator hand-over between the items of a gate (nchunks > 1)
  unsigned int* ctl;        // [0] next work item, [1] finished blocks; zero at launch, re-zeroed by the last block
  int* progress;            // [count] items finished per gate (nchunks > 1); zero at launch, re-zeroed by the last block
};

// first CMUX step of work item `ch` of a gate (ch == nchunks: n)
__device__ __forceinline__ int br_chunk_lo(const BrArgs& A, int ch) {
  if (ch < A.main_chunks) return ch * A.chunk_steps;
  return ch >= A.nchunks ? A.n : A.tail_start[ch - A.main_chunks];
}

struct CmuxArgs {
  const uint32_t* ct0;      // NULL => zero (plain external product)
  const uint32_t* ct1;      // [count][2][N]
  const double2* bsk_row;   // [2L][2][8][T]
  const double2* tw_tab;
  uint32_t* out;            // [count][2][N]
  uint32_t offset;
  Tw4 tw0;
};

// ---------------------------------------------------------------------------------------------
// butterflies.  Forward: (u, v) -> (u + w v, u - w v) in 6 FMA (second output as 2u - first).
// Inverse: (u, v) -> (u + v, (u - v) conj(w)).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bf_fwd(double2& u, double2& v, double wr, double wi) {
  double tr = fma(v.x, wr, u.x);
  tr = fma(-v.y, wi, tr);
  double ti = fma(v.x, wi, u.y);
  ti = fma(v.y, wr, ti);
  v.x = fma(2.0, u.x, -tr);
  v.y = fma(2.0, u.y, -ti);
  u.x = tr;
  u.y = ti;
}
__device__ __forceinline__ void bf_inv(double2& u, double2& v, double wr, double wi) {
  double dx = u.x - v.x, dy = u.y - v.y;
  u.x += v.x;
  u.y += v.y;
  v.x = fma(dx, wr, dy * wi);
  v.y = fma(dy, wr, -(dx * wi));
}

// Radix-8 block on 8 register-resident points x[0..7] = block elements at stride s.
// NST = 3: stages A,B,C; 2: B,C; 1: C.  Block 2i+1 uses -i * S(2m,2i) (= (wi, -wr)), etc.
template <int NST>
__device__ __forceinline__ void radix8_fwd(double2 (&x)[8], const double2& s0, const double2& s1, const double2& s2,
                                           const double2& s3) {
  if constexpr (NST >= 3) {
#pragma unroll
    for (int k = 0; k < 4; k++) bf_fwd(x[k], x[k + 4], s0.x, s0.y);
  }
  if constexpr (NST >= 2) {
    bf_fwd(x[0], x[2], s1.x, s1.y);
    bf_fwd(x[1], x[3], s1.x, s1.y);
    bf_fwd(x[4], x[6], s1.y, -s1.x);
    bf_fwd(x[5], x[7], s1.y, -s1.x);
  }
  bf_fwd(x[0], x[1], s2.x, s2.y);
  bf_fwd(x[2], x[3], s2.y, -s2.x);
  bf_fwd(x[4], x[5], s3.x, s3.y);
  bf_fwd(x[6], x[7], s3.y, -s3.x);
}
template <int NST>
__device__ __forceinline__ void radix8_inv(double2 (&x)[8], const double2& s0, const double2& s1, const double2& s2,
                                           const double2& s3) {
  bf_inv(x[0], x[1], s2.x, s2.y);
  bf_inv(x[2], x[3], s2.y, -s2.x);
  bf_inv(x[4], x[5], s3.x, s3.y);
  bf_inv(x[6], x[7], s3.y, -s3.x);
  if constexpr (NST >= 2) {
    bf_inv(x[0], x[2], s1.x, s1.y);
    bf_inv(x[1], x[3], s1.x, s1.y);
    bf_inv(x[4], x[6], s1.y, -s1.x);
    bf_inv(x[5], x[7], s1.y, -s1.x);
  }
  if constexpr (NST >= 3) {
#pragma unroll
    for (int k = 0; k < 4; k++) bf_inv(x[k], x[k + 4], s0.x, s0.y);
  }
}

// ---------------------------------------------------------------------------------------------
// Pass geometry for an M = 2^LOGM point transform with T = M/8 threads.
// Full pass k (0-based) works on elements at stride s_k = M >> 3(k+1); if LOGM is not a multiple
// of 3 a final partial pass (stride 1, 8 contiguous points) runs the last LOGM % 3 stages.
// ---------------------------------------------------------------------------------------------
template <int LOGM>
struct Geo {
  static constexpr int M = 1 << LOGM;
  static constexpr int T = M / 8;
  static constexpr int NFULL = LOGM / 3;
  static constexpr int REM = LOGM % 3;
  static constexpr int NPASS = NFULL + (REM ? 1 : 0);
  __host__ __device__ static constexpr int stride(int k) { return k < NFULL ? (M >> (3 * (k + 1))) : 1; }
  __host__ __device__ static constexpr int nstages(int k) { return k < NFULL ? 3 : REM; }
  // number of blocks entering pass k (twiddle table length), and table offset (in Tw4 units) for k >= 1
  __host__ __device__ static constexpr int blocks(int k) { return k < NFULL ? (1 << (3 * k)) : T; }
  __host__ __device__ static constexpr int tab_off(int k) {
    int o = 0;
    for (int q = 1; q < k; q++) o += blocks(q);
    return o;
  }
  __host__ __device__ static constexpr int tab_len() { return tab_off(NPASS); }
  // first element position of thread tau in pass k; element a sits at base + stride(k) * a
  __device__ __forceinline__ static int base(int k, int tau) {
    const int s = stride(k);
    return (tau / s) * (8 * s) + (tau % s);
  }
  __device__ __forceinline__ static int block_of(int k, int tau) { return tau / stride(k); }
};

// ---- mbarrier / bulk-copy (TMA) primitives ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// predicated 16-byte shared-memory load: v keeps its value where pred is false (no divergent branch)
__device__ __forceinline__ void lds_if(double2& v, const double2* p, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
      : "+d"(v.x), "+d"(v.y)
      : "r"(smem_u32(p)), "r"((int)pred)
      : "memory");
}

// predicated on (own != A) with the comparison inside the asm block: one ISETP + one predicated LDS/STS
template <int A>
__device__ __forceinline__ void lds_unless(double2& v, const double2* p, int own) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, %4;\n\t@q ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
               : "+d"(v.x), "+d"(v.y)
               : "r"(smem_u32(p)), "r"(own), "n"(A)
               : "memory");
}
template <int A>
__device__ __forceinline__ void sts_unless(double2* p, const double2& v, int own) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, %4;\n\t@q st.shared.v2.f64 [%0], {%1, %2};\n\t}" ::"r"(smem_u32(p)),
               "d"(v.x), "d"(v.y), "r"(own), "n"(A)
               : "memory");
}

__device__ __forceinline__ void sts_if(double2* p, const double2& v, bool pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q st.shared.v2.f64 [%0], {%1, %2};\n\t}" ::"r"(smem_u32(p)),
               "d"(v.x), "d"(v.y), "r"((int)pred)
               : "memory");
}

// 16-byte-slot swizzle: conflict-free for every pass stride used above.
__device__ __forceinline__ int swz(int p) { return p ^ ((p >> 3) & 7); }

// SINGLE = false: two exchange buffers used alternately (one block barrier per exchange).
// SINGLE = true : one exchange buffer; the write-after-read hazard is covered by an mbarrier on which every thread
//                 arrives (non-blocking) after its reads and waits before its next writes (normally long complete).
// NAMED = true : the transform's threads are one of several independent groups of T threads inside a bigger block
//                 (several gates per block): exchanges synchronise on the named barrier `bar_id` (1..15) instead of
//                 the whole block.
template <int LOGM, bool SINGLE = false, bool NAMED = false>
struct Fft {
  using G = Geo<LOGM>;
  int bar_id = 0;
  __device__ __forceinline__ void group_barrier() const {
    if constexpr (NAMED) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(G::T) : "memory");
    else __syncthreads();
  }
  uint64_t* rd_bar = nullptr;
  uint32_t rd_phase = 0;
  // Per-thread persistent state: twiddles of the last pass (unique per thread) and the ping-pong
  // parity of the exchange buffers.
  double2 tl0, tl1, tl2, tl3;
  int parity;
  double2* ex;            // [2][M] exchange buffers in shared memory
  const Tw4* tab;         // twiddle tables for passes >= 1
  int tau;

  // call from every thread, before a __syncthreads(); rd_bar_ must have been mbar_init'ed with count T
  __device__ __forceinline__ void init_single(uint64_t* rd_bar_) {
    rd_bar = rd_bar_;
    rd_phase = 0;
  }
  __device__ __forceinline__ void init(double2* ex_, const double2* tw_tab, int tau_) {
    ex = ex_;
    tab = reinterpret_cast<const Tw4*>(tw_tab);
    tau = tau_;
    parity = 0;
    const Tw4* e = tab + G::tab_off(G::NPASS - 1) + G::block_of(G::NPASS - 1, tau);
    tl0 = e->s[0]; tl1 = e->s[1]; tl2 = e->s[2]; tl3 = e->s[3];
  }

  struct NoHook { __device__ __forceinline__ void operator()() const {} };

  // hook() runs right after the block barrier (every thread has finished whatever preceded this exchange)
  template <int KW, int KR, class Hook>
  __device__ __forceinline__ void exchange(double2 (&x)[8], const Hook& hook) {
    double2* buf = ex;
    // threads that trade points in this exchange: stride(coarser pass) consecutive threads; inside one warp for the
    // exchanges next to the last pass (8 threads at N = 1024, 16 at N = 2048)
    constexpr bool WARP_LOCAL = br_warp_ex(LOGM + 1) && !SINGLE && !NAMED && (G::stride(KW < KR ? KW : KR) <= 32) && (G::T % 32 == 0);
    if constexpr (SINGLE) {
      mbar_wait(rd_bar, rd_phase);  // every thread has finished reading the previous exchange
      rd_phase ^= 1u;
    } else if constexpr (WARP_LOCAL) {
      buf += 2 * TFHE_BR_EXW * G::M;  // third buffer: never touched by a block-wide exchange
      __syncwarp();                   // the warp has finished reading its groups from the previous warp-local exchange
    } else {
      buf += (parity ? TFHE_BR_EXW * G::M : 0);
      parity ^= 1;
    }
    if (TFHE_BR_KO & 1) return;
    const int wb = G::base(KW, tau), rb = G::base(KR, tau);
    // Between two full passes exactly one of a thread's 8 points keeps both its owner and its register slot
    // (slot (tau / finer stride) % 8): it stays in its register, skipping 1/8 of the shared-memory traffic.
    // (mode 2: only if that slot is uniform over 8 consecutive lanes — the finer pass's stride is a multiple of 8 — so that
    // the skipped accesses of a 16-byte instruction form a whole quarter-warp, i.e. one of its four 128-byte wavefronts)
    constexpr bool KEEP = TFHE_BR_KEEP_OWN && (KW < G::NFULL) && (KR < G::NFULL) &&
                          (TFHE_BR_KEEP_OWN == 1 || G::stride(KW > KR ? KW : KR) % 8 == 0);
    const int own = KEEP ? ((tau / G::stride(KW > KR ? KW : KR)) & 7) : -1;
    if constexpr (KEEP) {
      sts_unless<0>(buf + swz(wb + G::stride(KW) * 0), x[0], own); sts_unless<1>(buf + swz(wb + G::stride(KW) * 1), x[1], own);
      sts_unless<2>(buf + swz(wb + G::stride(KW) * 2), x[2], own); sts_unless<3>(buf + swz(wb + G::stride(KW) * 3), x[3], own);
      sts_unless<4>(buf + swz(wb + G::stride(KW) * 4), x[4], own); sts_unless<5>(buf + swz(wb + G::stride(KW) * 5), x[5], own);
      sts_unless<6>(buf + swz(wb + G::stride(KW) * 6), x[6], own); sts_unless<7>(buf + swz(wb + G::stride(KW) * 7), x[7], own);
    } else {
#pragma unroll
      for (int a = 0; a < 8; a++) buf[swz(wb + G::stride(KW) * a)] = x[a];
    }
    if constexpr (WARP_LOCAL) __syncwarp();
    else if (!(TFHE_BR_KO & 4)) group_barrier();
    hook();
    if constexpr (KEEP) {  // predicated, never a branch
      lds_unless<0>(x[0], buf + swz(rb + G::stride(KR) * 0), own); lds_unless<1>(x[1], buf + swz(rb + G::stride(KR) * 1), own);
      lds_unless<2>(x[2], buf + swz(rb + G::stride(KR) * 2), own); lds_unless<3>(x[3], buf + swz(rb + G::stride(KR) * 3), own);
      lds_unless<4>(x[4], buf + swz(rb + G::stride(KR) * 4), own); lds_unless<5>(x[5], buf + swz(rb + G::stride(KR) * 5), own);
      lds_unless<6>(x[6], buf + swz(rb + G::stride(KR) * 6), own); lds_unless<7>(x[7], buf + swz(rb + G::stride(KR) * 7), own);
    } else {
#pragma unroll
      for (int a = 0; a < 8; a++) x[a] = buf[swz(rb + G::stride(KR) * a)];
    }
    if constexpr (SINGLE) mbar_arrive(rd_bar);
  }
  template <int KW, int KR>
  __device__ __forceinline__ void exchange(double2 (&x)[8]) { exchange<KW, KR>(x, NoHook()); }

  // Two independent transforms exchanged together: one block barrier for both, and twice the independent work
  // between barriers (the buffers are [2 parities][2 transforms][M]).  Not available in SINGLE mode.
  template <int KW, int KR>
  __device__ __forceinline__ void exchange2(double2 (&x)[8], double2 (&y)[8]) {
    static_assert(!SINGLE && TFHE_BR_EXW == 2, "paired exchange needs the double-width ping-pong buffers");
    double2* buf = ex + (parity ? 2 * G::M : 0);
    parity ^= 1;
    const int wb = G::base(KW, tau), rb = G::base(KR, tau);
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int p = swz(wb + G::stride(KW) * a);
      buf[p] = x[a];
      buf[G::M + p] = y[a];
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int p = swz(rb + G::stride(KR) * a);
      x[a] = buf[p];
      y[a] = buf[G::M + p];
    }
  }

  template <int K>
  __device__ __forceinline__ void fwd_pass(double2 (&x)[8], const Tw4& tw0) {
    if constexpr (K == 0) {
      radix8_fwd<3>(x, tw0.s[0], tw0.s[1], tw0.s[2], tw0.s[3]);
    } else if constexpr (K == G::NPASS - 1 && !TFHE_BR_TL_RELOAD) {
      radix8_fwd<G::nstages(K)>(x, tl0, tl1, tl2, tl3);
    } else if constexpr (K == G::NPASS - 1) {  // last-pass twiddles re-read from L1 instead of living in 16 registers
      const Tw4* e = tab + G::tab_off(K) + G::block_of(K, tau);
      double2 s0 = __ldg(&e->s[0]), s1 = __ldg(&e->s[1]), s2 = __ldg(&e->s[2]), s3 = __ldg(&e->s[3]);
      radix8_fwd<G::nstages(K)>(x, s0, s1, s2, s3);
    } else {
#if TFHE_BR_TW_CONST
      if constexpr (K == 1) {
        const Tw4& e = c_tw_pass1[LOGM - 8][G::block_of(K, tau)];
        radix8_fwd<3>(x, e.s[0], e.s[1], e.s[2], e.s[3]);
        return;
      }
#endif
      const Tw4* e = tab + G::tab_off(K) + G::block_of(K, tau);
      double2 s0 = __ldg(&e->s[0]), s1 = __ldg(&e->s[1]), s2 = __ldg(&e->s[2]), s3 = __ldg(&e->s[3]);
      radix8_fwd<3>(x, s0, s1, s2, s3);
    }
  }
  template <int K>
  __device__ __forceinline__ void inv_pass(double2 (&x)[8], const Tw4& tw0) {
    if constexpr (K == 0) {
      radix8_inv<3>(x, tw0.s[0], tw0.s[1], tw0.s[2], tw0.s[3]);
    } else if constexpr (K == G::NPASS - 1 && !TFHE_BR_TL_RELOAD) {
      radix8_inv<G::nstages(K)>(x, tl0, tl1, tl2, tl3);
    } else if constexpr (K == G::NPASS - 1) {
      const Tw4* e = tab + G::tab_off(K) + G::block_of(K, tau);
      double2 s0 = __ldg(&e->s[0]), s1 = __ldg(&e->s[1]), s2 = __ldg(&e->s[2]), s3 = __ldg(&e->s[3]);
      radix8_inv<G::nstages(K)>(x, s0, s1, s2, s3);
    } else {
#if TFHE_BR_TW_CONST
      if constexpr (K == 1) {
        const Tw4& e = c_tw_pass1[LOGM - 8][G::block_of(K, tau)];
        radix8_inv<3>(x, e.s[0], e.s[1], e.s[2], e.s[3]);
        return;
      }
#endif
      const Tw4* e = tab + G::tab_off(K) + G::block_of(K, tau);
      double2 s0 = __ldg(&e->s[0]), s1 = __ldg(&e->s[1]), s2 = __ldg(&e->s[2]), s3 = __ldg(&e->s[3]);
      radix8_inv<3>(x, s0, s1, s2, s3);
    }
  }

  // in: x[a] = z[tau + T a] (folded coefficients); out: x[e] = spectrum at position 8 tau + e
  // When the final partial pass is a single radix-2 stage (LOGM % 3 == 1, i.e. N = 2048), the two points of every
  // butterfly sit in neighbouring lanes (stride-2 layout of the previous pass): instead of a third shared-memory
  // exchange, lanes tau and tau^1 trade half of their points by warp shuffle and each finishes 4 butterflies.
  static constexpr bool SHFL_LAST = TFHE_BR_SHFL_LAST && (G::REM == 1) && (G::NFULL >= 1);
  __device__ __forceinline__ void swap_with_neighbour(double2 (&x)[8]) const {
    const bool u = tau & 1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double sx = u ? x[k].x : x[k + 4].x, sy = u ? x[k].y : x[k + 4].y;
      const double rx = __shfl_xor_sync(0xffffffffu, sx, 1), ry = __shfl_xor_sync(0xffffffffu, sy, 1);
      x[k].x = u ? rx : x[k].x; x[k].y = u ? ry : x[k].y;
      x[k + 4].x = u ? x[k + 4].x : rx; x[k + 4].y = u ? x[k + 4].y : ry;
    }
  }
  // in: x[a] = point at (tau/2)*16 + tau%2 + 2a (layout of the last full pass); out: x[e] = point at 8 tau + e
  __device__ __forceinline__ void last_stage_fwd(double2 (&x)[8]) const {
    swap_with_neighbour(x);  // now (x[k], x[k+4]) = (even, odd) point of butterfly 4 tau + k
    bf_fwd(x[0], x[4], tl2.x, tl2.y);
    bf_fwd(x[1], x[5], tl2.y, -tl2.x);
    bf_fwd(x[2], x[6], tl3.x, tl3.y);
    bf_fwd(x[3], x[7], tl3.y, -tl3.x);
    const double2 y1 = x[4], y2 = x[1], y3 = x[5], y4 = x[2], y5 = x[6], y6 = x[3];  // to natural order e = 2k + {0,1}
    x[1] = y1; x[2] = y2; x[3] = y3; x[4] = y4; x[5] = y5; x[6] = y6;
  }
  __device__ __forceinline__ void last_stage_inv(double2 (&x)[8]) const {
    const double2 y1 = x[1], y2 = x[2], y3 = x[3], y4 = x[4], y5 = x[5], y6 = x[6];
    x[4] = y1; x[1] = y2; x[5] = y3; x[2] = y4; x[6] = y5; x[3] = y6;
    bf_inv(x[0], x[4], tl2.x, tl2.y);
    bf_inv(x[1], x[5], tl2.y, -tl2.x);
    bf_inv(x[2], x[6], tl3.x, tl3.y);
    bf_inv(x[3], x[7], tl3.y, -tl3.x);
    swap_with_neighbour(x);
  }

  template <class Hook>
  __device__ __forceinline__ void forward(double2 (&x)[8], const Tw4& tw0, const Hook& hook) {
    fwd_pass<0>(x, tw0);
    if constexpr (G::NPASS > 1) { exchange<0, 1>(x, hook); fwd_pass<1>(x, tw0); }
    if constexpr (G::NPASS > 2) { exchange<1, 2>(x); fwd_pass<2>(x, tw0); }
    if constexpr (G::NPASS > 3) {
      if constexpr (SHFL_LAST) last_stage_fwd(x);
      else { exchange<2, 3>(x); fwd_pass<3>(x, tw0); }
    }
  }
  __device__ __forceinline__ void forward(double2 (&x)[8], const Tw4& tw0) { forward(x, tw0, NoHook()); }
  // same with a second hook after the barrier of the LAST exchange (N = 1024: h1 -> pass 1 -> h2 -> pass 2)
  template <class Hook1, class Hook2>
  __device__ __forceinline__ void forward_hooks(double2 (&x)[8], const Tw4& tw0, const Hook1& h1, const Hook2& h2) {
    fwd_pass<0>(x, tw0);
    if constexpr (G::NPASS == 3) {
      exchange<0, 1>(x, h1); fwd_pass<1>(x, tw0);
      exchange<1, 2>(x, h2); fwd_pass<2>(x, tw0);
    } else {
      forward_rest(x, tw0, h1);
      h2();
    }
  }
  template <class Hook>
  __device__ __forceinline__ void forward_rest(double2 (&x)[8], const Tw4& tw0, const Hook& hook) {  // forward() minus pass 0
    if constexpr (G::NPASS > 1) { exchange<0, 1>(x, hook); fwd_pass<1>(x, tw0); }
    if constexpr (G::NPASS > 2) { exchange<1, 2>(x); fwd_pass<2>(x, tw0); }
    if constexpr (G::NPASS > 3) {
      if constexpr (SHFL_LAST) last_stage_fwd(x);
      else { exchange<2, 3>(x); fwd_pass<3>(x, tw0); }
    }
  }
  // forward() split in two so that a caller can put long-latency loads in flight before the last register pass
  __device__ __forceinline__ void forward_head(double2 (&x)[8], const Tw4& tw0) {
    fwd_pass<0>(x, tw0);
    if constexpr (G::NPASS > 2) { exchange<0, 1>(x); fwd_pass<1>(x, tw0); }
    if constexpr (G::NPASS > 3) { exchange<1, 2>(x); fwd_pass<2>(x, tw0); }
    if constexpr (G::NPASS > 1) exchange<G::NPASS - 2, G::NPASS - 1>(x);
  }
  __device__ __forceinline__ void forward_tail(double2 (&x)[8], const Tw4& tw0) {
    if constexpr (G::NPASS > 1) fwd_pass<G::NPASS - 1>(x, tw0);
  }
  // exact inverse of forward() up to the factor M (folded into the bootstrapping key); hook after the FIRST barrier
  template <class Hook>
  __device__ __forceinline__ void inverse(double2 (&x)[8], const Tw4& tw0, const Hook& hook) {
    if constexpr (G::NPASS > 3) {
      if constexpr (SHFL_LAST) { last_stage_inv(x); inv_pass<2>(x, tw0); exchange<2, 1>(x, hook); }
      else { inv_pass<3>(x, tw0); exchange<3, 2>(x, hook); inv_pass<2>(x, tw0); exchange<2, 1>(x); }
      inv_pass<1>(x, tw0); exchange<1, 0>(x);
    } else if constexpr (G::NPASS > 2) {
      inv_pass<2>(x, tw0); exchange<2, 1>(x, hook);
      inv_pass<1>(x, tw0); exchange<1, 0>(x);
    } else if constexpr (G::NPASS > 1) {
      inv_pass<1>(x, tw0); exchange<1, 0>(x, hook);
    }
    inv_pass<0>(x, tw0);
  }
  __device__ __forceinline__ void inverse(double2 (&x)[8], const Tw4& tw0) { inverse(x, tw0, NoHook()); }

  // paired versions: the same passes on two independent arrays, exchanges shared
  __device__ __forceinline__ void forward2(double2 (&x)[8], double2 (&y)[8], const Tw4& tw0) {
    fwd_pass<0>(x, tw0); fwd_pass<0>(y, tw0);
    if constexpr (G::NPASS > 1) { exchange2<0, 1>(x, y); fwd_pass<1>(x, tw0); fwd_pass<1>(y, tw0); }
    if constexpr (G::NPASS > 2) { exchange2<1, 2>(x, y); fwd_pass<2>(x, tw0); fwd_pass<2>(y, tw0); }
    if constexpr (G::NPASS > 3) {
      if constexpr (SHFL_LAST) { last_stage_fwd(x); last_stage_fwd(y); }
      else { exchange2<2, 3>(x, y); fwd_pass<3>(x, tw0); fwd_pass<3>(y, tw0); }
    }
  }
  __device__ __forceinline__ void inverse2(double2 (&x)[8], double2 (&y)[8], const Tw4& tw0) {
    if constexpr (G::NPASS > 3) {
      if constexpr (SHFL_LAST) { last_stage_inv(x); last_stage_inv(y); }
      else { inv_pass<3>(x, tw0); inv_pass<3>(y, tw0); exchange2<3, 2>(x, y); }
    }
    if constexpr (G::NPASS > 2) { inv_pass<2>(x, tw0); inv_pass<2>(y, tw0); exchange2<2, 1>(x, y); }
    if constexpr (G::NPASS > 1) { inv_pass<1>(x, tw0); inv_pass<1>(y, tw0); exchange2<1, 0>(x, y); }
    inv_pass<0>(x, tw0); inv_pass<0>(y, tw0);
  }
};

// ---------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------
// exact int -> double for 0 <= field < 2^32 minus a constant bias, without I2F:
// bits(2^52 + field) then one DADD.
__device__ __forceinline__ double field_to_double(uint32_t field, double bias) {
  return __hiloint2double(0x43300000, (int)field) - bias;
}

// Gadget digit (poly/decomposer.go:55-66) of the level whose field starts at bit `sh` of w = coefficient + offset,
// returned SCALED by 2^sh: ((w >> sh) & (Bg-1)) - Bg/2, times 2^sh.  The field is masked IN PLACE (no shift) into the
// low mantissa word of 2^52, so a digit costs one LOP3 + one DADD.  Transforms are linear and powers of two are exact,
// so the spectrum is the unscaled one times 2^sh; the bootstrapping-key rows of that level are stored scaled by 2^-sh
// (bsk_repack_kernel) and every product x * k — hence every rounding — is bit-identical to the unscaled computation.
template <int BGBIT>
__device__ __forceinline__ double digit_scaled(uint32_t w, int sh) {
  constexpr uint32_t MASK0 = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  const double v = __hiloint2double(0x43300000, (int)(w & (MASK0 << sh)));
  const double bias = __hiloint2double(0x43300000, (int)(1u << (BGBIT - 1 + sh)));  // 2^52 + (Bg/2) 2^sh
  return v - bias;
}
// scale applied to the key rows that multiply the digits of level lvl (rows lvl and L + lvl of every row-set)
__host__ __device__ inline double digit_key_scale(int bgbit, int lvl) {
  double s = 1.0;
  for (int k = 0; k < 32 - (lvl + 1) * bgbit; k++) s *= 0.5;
  return s;
}

// Spectrum-domain result -> torus word.  poly/fourier_transform.go:88-125: round(x - 2^32 round(x / 2^32))
// then uint32(int64(.)).  SMALL: |y| < 2^51 guaranteed by the parameter set, so one magic-number add
// yields round-to-nearest(y) mod 2^32 (ties cannot occur where the result is exact).
template <bool SMALL>
__device__ __forceinline__ uint32_t to_torus(double y) {
  if (!SMALL) {
    double q = rint(y * (1.0 / 4294967296.0));
    y = fma(-4294967296.0, q, y);
  }
  return (uint32_t)__double2loint(y + 6755399441055744.0);
}

// value of (X^k * P)[j] given idx = (j - k) mod 2N: P[idx] or ~P[idx - N]  (buffer_methods.go:133-164;
// the wrap-around "negation" is 0xFFFFFFFF - a, not -a)
template <int N>
__device__ __forceinline__ uint32_t rot_read(const uint32_t* P, int idx) {
  uint32_t v = P[idx & (N - 1)];
  return (idx & N) ? ~v : v;
}

// ---------------------------------------------------------------------------------------------
// One CMUX step on the shared-memory accumulator: acc += BK (x) (X^at * acc - acc).
// ---------------------------------------------------------------------------------------------
// key-row fetch policies for the MAC: straight LDG (LSU pipe) or texture fetch (TEX pipe)
// Device layout of one bootstrapping-key row-set (2L rows x {A,B} x M spectrum values, double2 units).
// TFHE_BR_KEY256 = 0 (default): [row][{A,B}][e][tau] — every warp-wide 16-byte load covers 512 contiguous bytes.
// TFHE_BR_KEY256 = 1: the A and the B value a thread multiplies one spectrum point with are adjacent (32 bytes), so a
// multiply-accumulate fetches them with ONE 256-bit load (LDG.E.ENL2.256): half the load instructions for the same
// bytes.  Measured (profiles/r02_experiments.md): no gain for the throughput kernel (its time follows LSU wavefronts, not
// instructions) and the latency kernels' 16-byte cp.async pieces then use half of every 32-byte sector (single gate
// 2.5 -> 3.7 ms), so it stays off.
#ifndef TFHE_BR_KEY256
#define TFHE_BR_KEY256 0
#endif
template <int T>
__host__ __device__ constexpr int key_pos(int r, int ab, int e, int tau) {
#if TFHE_BR_KEY256
  return ((r * 8 + e) * T + tau) * 2 + ab;
#else
  return ((r * 2 + ab) * 8 + e) * T + tau;
#endif
}
#ifndef TFHE_BR_KEYLD
#define TFHE_BR_KEYLD 0   // how key rows are loaded: 0 = ld.global.nc (L1-allocating), 1 = ld.global.cg (L2 only), 2 = ld.global.nc.L1::no_allocate (same as 0); experiments: 3 = nc.L1::evict_first, 4 = nc.L1::evict_last, 5 = nc.L2::256B, 6 = cg.L2::256B, 7 = cv
#endif
#ifndef TFHE_BR_KEYLD_BIG
#define TFHE_BR_KEYLD_BIG 1   // policy from ring size TFHE_BR_KEYLD_CG_LOGN on
#endif
#ifndef TFHE_BR_KEYLD_CG_LOGN
#define TFHE_BR_KEYLD_CG_LOGN 11  // ... and ld.global.cg from this ring size on.  Measured (r02_experiments.md): cg is 2 % SLOWER at N = 1024 (128-bit, 80-bit, Uint1, Uint3: 3 %) and 2.2 % FASTER at N = 2048 (Uint5 39.31 -> 38.44 ms, Uint4 30.14 -> 29.49)
#endif
template <int POLICY>
struct KeyLdgT {
  const double2* __restrict__ p;
  __device__ __forceinline__ double2 operator()(int idx) const {
    if (TFHE_BR_KO & 2) return make_double2(1e-9 * idx, 2e-9 * idx);
    if constexpr (POLICY == 1) {
      return __ldcg(p + idx);
    } else if constexpr (POLICY >= 2 && POLICY <= 7) {  // experiment policies (see TFHE_BR_KEYLD)
      double2 v;
      if constexpr (POLICY == 2) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + idx));
      if constexpr (POLICY == 3) asm volatile("ld.global.nc.L1::evict_first.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + idx));
      if constexpr (POLICY == 4) asm volatile("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + idx));
      if constexpr (POLICY == 5) asm volatile("ld.global.nc.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + idx));
      if constexpr (POLICY == 6) asm volatile("ld.global.cg.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + idx));
      if constexpr (POLICY == 7) asm volatile("ld.global.cv.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + idx));
      return v;
    } else {
      return __ldg(p + idx);
    }
  }
  // the two key values (A row, B row) for spectrum slot e of row r
  template <int T>
  __device__ __forceinline__ void pair(int r, int e, int tau, double2& ka, double2& kb) const {
#if TFHE_BR_KEY256
    if (TFHE_BR_KO & 2) { ka = make_double2(1e-9 * e, 2e-9 * r); kb = ka; return; }
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(ka.x), "=d"(ka.y), "=d"(kb.x), "=d"(kb.y)
                 : "l"(p + key_pos<T>(r, 0, e, tau)));
#else
    ka = (*this)(key_pos<T>(r, 0, e, tau));
    kb = (*this)(key_pos<T>(r, 1, e, tau));
#endif
  }
};
using KeyLdg = KeyLdgT<TFHE_BR_KEYLD>;
template <int LOGN>
using KeyLdgFor = KeyLdgT<(LOGN >= TFHE_BR_KEYLD_CG_LOGN) ? TFHE_BR_KEYLD_BIG : TFHE_BR_KEYLD>;  // the throughput kernel's policy per ring size
struct KeyTex {
  cudaTextureObject_t tex;
  int base;  // in double2 units
  __device__ __forceinline__ double2 operator()(int idx) const {
    const uint4 v = tex1Dfetch<uint4>(tex, base + idx);
    return make_double2(__hiloint2double((int)v.y, (int)v.x), __hiloint2double((int)v.w, (int)v.z));
  }
  template <int T>
  __device__ __forceinline__ void pair(int r, int e, int tau, double2& ka, double2& kb) const {
    ka = (*this)(key_pos<T>(r, 0, e, tau));
    kb = (*this)(key_pos<T>(r, 1, e, tau));
  }
};

// ---- shared prologue / epilogue pieces of every blind-rotate kernel ------------------------------------------------
// test vector of gate g: CloudKey.BlindRotateTestvec, one LUT for all, a LUT per ciphertext, or an index into a LUT table
template <int N>
__device__ __forceinline__ const uint32_t* br_testvec(const BrArgs& A, long long g) {
  if (!A.luts) return A.testvec;
  const long long k = A.lut_index ? (long long)A.lut_index[g] : (A.nluts == 1 ? 0 : g);
  return A.luts + k * (2 * N);
}
// mod switch (evaluator.go:116,122): a~ = ((a + 2^(30-NBIT)) mod 2^32) >> (31-NBIT).  lk > 0 (many-LUT bootstraps, no
// reference counterpart): the switch goes to 2N / 2^lk levels, scaled back up, so that a~ is a multiple of 2^lk.
template <int LOGN>
__device__ __forceinline__ int br_modswitch(uint32_t a, int lk) {
  return (int)(((a + (1u << (30 - LOGN + lk))) >> (31 - LOGN + lk)) << lk);
}
// b~ = 2N - ((int64(b) + 2^(30-NBIT)) >> (31-NBIT))  (the sum is taken in 64 bits: b~ = 0 when it carries)
template <int LOGN>
__device__ __forceinline__ int br_btilde(uint32_t b, int lk) {
  const unsigned long long bb = (unsigned long long)b + (1ull << (30 - LOGN + lk));
  return (int)((2 * (1 << LOGN) - (int)((bb >> (31 - LOGN + lk)) << lk)) & (2 * (1 << LOGN) - 1));
}
// epilogue: TRLWE (mode 0), sample extract at 0 (mode 1; trlwe/trlwe_ops.go:10-21: out[0] = A[0], out[i] = ~A[N-i],
// out[N] = B[0]) or at indices 0..K-1 (mode 2; trlwe.SampleExtractIndex trlwe/trlwe.go:114-128: out[i] = A[k-i] for
// i <= k, ~A[N+k-i] above, out[N] = B[k]).  PA / PB: the accumulator polynomials; tid / nt: this thread among the gate's.
template <int N>
__device__ __forceinline__ void br_write_output(const BrArgs& A, long long g, const uint32_t* PA, const uint32_t* PB, int tid, int nt) {
  if (A.out_mode == 0) {
    uint32_t* o = A.out + g * (2 * N);
    for (int j = tid; j < N; j += nt) { o[j] = PA[j]; o[N + j] = PB[j]; }
  } else if (A.out_mode == 1) {
    uint32_t* o = A.out + g * (N + 1);
    for (int j = tid; j < N; j += nt) o[j] = (j == 0) ? PA[0] : ~PA[N - j];
    if (tid == 0) o[N] = PB[0];
  } else {
    for (int k = 0; k < A.extract_k; k++) {
      uint32_t* o = A.out + (g * A.extract_k + k) * (N + 1);
      for (int j = tid; j < N; j += nt) o[j] = (j <= k) ? PA[k - j] : ~PA[N + k - j];
      if (tid == 0) o[N] = PB[k];
    }
  }
}

// Accumulator TRLWE in shared memory: [2 polynomials][EXT * N] words.  EXT = 1: the N coefficients.  EXT = 2: followed
// by their complements, so that (X^k P)[j] (buffer_methods.go:133-164: P[idx] or ~P[idx - N], idx = (j - k) mod 2N) is
// one load at idx.  EXT = 3: followed by the coefficients again, so that idx + offsets < 3N needs no wrap either.
#ifndef TFHE_BR_ACC_EXT
#define TFHE_BR_ACC_EXT 1
#endif
template <int LOGN>
struct AccBuf {
  static constexpr int N = 1 << LOGN, EXT = TFHE_BR_ACC_EXT, STRIDE = EXT * N;
  __device__ __forceinline__ static void store(uint32_t* P, int j, uint32_t v) {
    P[j] = v;
    if constexpr (EXT >= 2) P[N + j] = ~v;
    if constexpr (EXT >= 3) P[2 * N + j] = v;
  }
};

// TFHE_BR_KEY_EARLY: the 16 key values of a multiply-accumulate are requested (pinned by volatile asm) 1 = right after the
// barrier of the transform's first exchange, 2 = right after the barrier of its last exchange, instead of wherever ptxas
// places the loads (mostly inside the last register pass, ~150 cycles before their use: ncu shows 5 % of all stall samples
// on the first two DFMAs of the multiply-accumulate, waiting on L2).  Costs 64 registers for the time in between.
#ifndef TFHE_BR_KEY_EARLY
#define TFHE_BR_KEY_EARLY 0
#endif
template <int LOGN, int L, int BGBIT, bool SMALL, class Key>
__device__ __forceinline__ void cmux_rotate_step(uint32_t* acc, Fft<LOGN - 1, false>& fft, const Key bk, int at,
                                                 uint32_t offset, const Tw4& tw0) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  using AB = AccBuf<LOGN>;
  const int tau = fft.tau;
  double2 accA[8], accB[8];
#pragma unroll
  for (int e = 0; e < 8; e++) { accA[e] = make_double2(0.0, 0.0); accB[e] = make_double2(0.0, 0.0); }

  // spectrum of one digit polynomial times key row r, accumulated into both output spectra
  auto mac = [&](const double2 (&x)[8], int r) {
#if TFHE_BR_PF_L1
    if constexpr (!std::is_same<Key, KeyTex>::value) {  // rows of digit r + d (contiguous into the next step's row-set)
      const char* pf = reinterpret_cast<const char*>(bk.p + (r + TFHE_BR_PF_L1) * 2 * M) + tau * 128;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + T * 128));
    }
#endif
#pragma unroll
    for (int e = 0; e < 8; e++) {
      double2 ka, kb;
      bk.template pair<T>(r, e, tau, ka, kb);
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

  TFHE_UNROLL(TFHE_BR_UNROLL_POLY)
  for (int poly = 0; poly < 2; poly++) {
    const uint32_t* P = acc + poly * AB::STRIDE;
    // decomposition input w = (X^at * P - P)[j] + offset for this thread's 16 coefficients j = tau + T a' (a' < 16);
    // rotated index idx = (j - at) mod 2N = ib + T a' walks N consecutive multiples of T: it wraps past N exactly once.
    uint32_t dre[8], dim[8];
    const int ib = (tau - at) & (2 * N - 1);
    if constexpr (AB::EXT == 3) {
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        dre[a] = P[ib + T * a] - P[j] + offset;
        dim[a] = P[ib + T * a + M] - P[j + M] + offset;
      }
    } else if constexpr (AB::EXT == 2) {
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        dre[a] = P[(ib + T * a) & (2 * N - 1)] - P[j] + offset;
        dim[a] = P[(ib + T * a + M) & (2 * N - 1)] - P[j + M] + offset;
      }
    } else {
      const int p0 = ib & (N - 1);
      const uint32_t s0 = 0u - (uint32_t)((ib >> LOGN) & 1);  // complement mask of a' = 0
      const int wrap = (N - p0 + T - 1) / T;                  // first a' whose index wraps (1..16): the mask flips there
#pragma unroll
      for (int a = 0; a < 8; a++) {
        const int j = tau + T * a;
        const uint32_t mre = s0 ^ (uint32_t)((wrap - 1 - a) >> 31), mim = s0 ^ (uint32_t)((wrap - 9 - a) >> 31);
        dre[a] = (P[(p0 + T * a) & (N - 1)] ^ mre) - P[j] + offset;
        dim[a] = (P[(p0 + T * a + M) & (N - 1)] ^ mim) - P[j + M] + offset;
      }
    }
    TFHE_UNROLL(TFHE_BR_UNROLL_LVL)
    for (int lvl = 0; lvl < L; lvl++) {
      double2 x[8];
      const int sh = 32 - (lvl + 1) * BGBIT;
#pragma unroll
      for (int a = 0; a < 8; a++) {
        x[a].x = digit_scaled<BGBIT>(dre[a], sh);
        x[a].y = digit_scaled<BGBIT>(dim[a], sh);
      }
#if TFHE_BR_KEY_EARLY
      if constexpr (!std::is_same<Key, KeyTex>::value) {
        double2 kA[8], kB[8];
        const int r = poly * L + lvl;
        auto load_keys = [&]() {
#pragma unroll
          for (int e = 0; e < 8; e++) {
            asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(kA[e].x), "=d"(kA[e].y) : "l"(bk.p + key_pos<T>(r, 0, e, tau)));
            asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(kB[e].x), "=d"(kB[e].y) : "l"(bk.p + key_pos<T>(r, 1, e, tau)));
          }
        };
        typename Fft<LOGN - 1, false>::NoHook nh;
        if (TFHE_BR_KEY_EARLY == 1) fft.forward_hooks(x, tw0, load_keys, nh);
        else fft.forward_hooks(x, tw0, nh, load_keys);
#pragma unroll
        for (int e = 0; e < 8; e++) {
          accA[e].x = fma(x[e].x, kA[e].x, accA[e].x);
          accA[e].x = fma(-x[e].y, kA[e].y, accA[e].x);
          accA[e].y = fma(x[e].x, kA[e].y, accA[e].y);
          accA[e].y = fma(x[e].y, kA[e].x, accA[e].y);
          accB[e].x = fma(x[e].x, kB[e].x, accB[e].x);
          accB[e].x = fma(-x[e].y, kB[e].y, accB[e].x);
          accB[e].y = fma(x[e].x, kB[e].y, accB[e].y);
          accB[e].y = fma(x[e].y, kB[e].x, accB[e].y);
        }
        continue;
      }
#endif
      fft.forward(x, tw0);
      mac(x, poly * L + lvl);
    }
  }
  fft.inverse(accA, tw0);
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
    AB::store(acc, j, acc[j] + to_torus<SMALL>(accA[a].x));
    AB::store(acc, j + M, acc[j + M] + to_torus<SMALL>(accA[a].y));
  }
  fft.inverse(accB, tw0);
  uint32_t* PB = acc + AB::STRIDE;
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
    AB::store(PB, j, PB[j] + to_torus<SMALL>(accB[a].x));
    AB::store(PB, j + M, PB[j + M] + to_torus<SMALL>(accB[a].y));
  }
}

template <int LOGN>
constexpr size_t br_smem_bytes(int n) {
  return (size_t)8 * TFHE_BR_ACC_EXT * (1 << LOGN) /*acc*/ + (size_t)br_nbuf(LOGN) * TFHE_BR_EXW * (1 << (LOGN - 1)) * 16 /*exchange [NBUF][EXW][M]*/ +
         (size_t)(((n + 1) * 2 + 15) / 16 * 16) /*abar*/;
}

// Bulk L2 prefetch (TMA unit, no registers, no LSU wavefronts): cp.async.bulk.prefetch.L2 of `bytes` (multiple of 16) at p.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#ifndef TFHE_BR_PF_MAX_LOGN
#define TFHE_BR_PF_MAX_LOGN 10  // largest ring size (log2) with the bulk L2 prefetch of the next chunk's key rows
#endif
#ifndef TFHE_BR_L2_PREFETCH
#define TFHE_BR_L2_PREFETCH 1   // work items request the key rows of the NEXT chunk (and, at kernel start, of the first) into L2 (N <= 1024: +0.13 % at 128-bit; N = 2048: -0.8 %, off there)
#endif

// acquire / release on the per-gate progress words of the work-item hand-over
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// The throughput kernel: PERSISTENT blocks of T = N/16 threads (as many as are resident at once) take work items from
// a global counter.  A work item is `chunk_steps` consecutive CMUX steps of one gate; the items of one gate run in
// order (possibly on different SMs), handing the accumulator TRLWE over through `scratch` in global memory (8 KiB per
// item at N = 1024, L2-resident).  Items are numbered chunk-major — item = chunk * count + gate — so the predecessor
// of an item was handed out `count` items earlier and is long finished when the batch exceeds the resident blocks; the
// acquire on `progress[gate]` makes that a guarantee instead of an assumption.  With nchunks = 1 a work item is a
// whole gate and nothing is handed over.  Why: a 4096-gate batch is 6.9 waves of 592 resident gates; whole-gate
// scheduling leaves SMs part-empty for the last wave, 50-step items leave them part-empty for 1/14 of a wave.
// ctl[0] (next item), ctl[1] (finished blocks) and progress[] must be zero at launch; the last block to finish
// re-zeroes them, so back-to-back launches need no memset.
// ---------------------------------------------------------------------------------------------
// TFHE_BR_LB_THREADS / TFHE_BR_LB_BLOCKS: experiment knob — declare looser launch bounds than the real block size to
// steer ptxas to a register budget between the (64,4) -> 255 and (64,5) -> 168 choices, e.g. (160,2) -> ~200.
#if defined(TFHE_BR_MAXNREG)   // experiment knob: an explicit register cap for every instance of the throughput kernel
#define TFHE_BR_BOUNDS(T, MINB) __maxnreg__(TFHE_BR_MAXNREG)
#elif defined(TFHE_BR_LB_THREADS)
#define TFHE_BR_BOUNDS(T, MINB) __launch_bounds__(TFHE_BR_LB_THREADS, TFHE_BR_LB_BLOCKS)
#else
#define TFHE_BR_BOUNDS(T, MINB) __launch_bounds__(T, MINB)
#endif
template <int LOGN, int L, int BGBIT, bool SMALL, int MINB, bool TEX = false>
__global__ void TFHE_BR_BOUNDS((1 << (LOGN - 4)), MINB) blind_rotate_kernel(const BrArgs A) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  using AB = AccBuf<LOGN>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                                 // [2][EXT * N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * AB::STRIDE);                   // [NBUF][EXW][M]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * AB::STRIDE + 16 * br_nbuf(LOGN) * TFHE_BR_EXW * M);
  __shared__ unsigned int s_item;
  const int tau = threadIdx.x;
  const int n = A.n;
  const unsigned long long total = (unsigned long long)A.count * (unsigned)A.nchunks;
  Fft<LOGN - 1> fft;
  fft.init(ex, A.tw_tab, tau);
  const size_t row_stride = (size_t)2 * L * 2 * M;
#ifdef TFHE_BR_STAGGER  // experiment: start the co-resident blocks of an SM a fraction of a step apart
  __nanosleep((blockIdx.x / 148u) * TFHE_BR_STAGGER);
#endif
#ifdef TFHE_BR_STAGGER_BLOCK  // experiment: every block of the grid starts TFHE_BR_STAGGER_BLOCK ns after its predecessor
  __nanosleep(blockIdx.x * TFHE_BR_STAGGER_BLOCK);
#endif
#if TFHE_BR_L2_PREFETCH
  // The key rows of a chunk are first touched by whichever blocks reach it first, and those wait out an HBM round trip
  // per digit; at kernel start (cold L2) that is EVERY block.  A chunk's rows are therefore requested into L2 ahead of
  // time, in slices, by the bulk-copy unit: chunk 0 by all blocks at entry, chunk c + 1 by the first items of chunk c.
  constexpr int PF_SLICES = 256;
  auto prefetch_chunk = [&](int ch, unsigned slice) {
    const size_t lo = (size_t)br_chunk_lo(A, ch), hi = (size_t)br_chunk_lo(A, ch + 1);
    if (lo >= hi) return;
    const size_t bytes = (hi - lo) * row_stride * sizeof(double2), per = (bytes / PF_SLICES + 15) / 16 * 16;
    const size_t off = (size_t)slice * per;
    if (off < bytes)
      prefetch_l2_bulk(reinterpret_cast<const char*>(A.bsk + lo * row_stride) + off, (uint32_t)min(per, bytes - off));
  };
  constexpr bool PF_ON = LOGN <= TFHE_BR_PF_MAX_LOGN;
  if (PF_ON && tau == 0 && blockIdx.x < PF_SLICES) prefetch_chunk(0, blockIdx.x);
#endif

  for (;;) {
    if (tau == 0) s_item = atomicAdd(&A.ctl[0], 1u);
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= total) break;
    const int chunk = (int)(item / (unsigned long long)A.count);
    const long long g = (long long)(item - (unsigned long long)chunk * (unsigned long long)A.count);
    const int i0 = br_chunk_lo(A, chunk), i1 = br_chunk_lo(A, chunk + 1);
#if TFHE_BR_L2_PREFETCH
    if (PF_ON && tau == 0 && g < PF_SLICES && chunk + 1 < A.nchunks) prefetch_chunk(chunk + 1, (unsigned)g);
#endif
    const uint32_t* __restrict__ ct = A.ct_in + g * (n + 1);
    // mod switch (evaluator.go:116,122): a~_i = ((a_i + 2^(30-NBIT)) mod 2^32) >> (31-NBIT)
    for (int i = i0 + tau; i < i1; i += T) abar[i - i0] = (unsigned short)br_modswitch<LOGN>(ct[i], A.ms_log2k);
    if (chunk == 0) {  // acc = X^btil * testvec
      const int btil = br_btilde<LOGN>(ct[n], A.ms_log2k);
      const uint32_t* __restrict__ tv = br_testvec<N>(A, g);
      for (int j = tau; j < N; j += T) {
        const int idx = (j - btil) & (2 * N - 1);
        const uint32_t va = tv[idx & (N - 1)], vb = tv[N + (idx & (N - 1))];
        AB::store(acc, j, (idx & N) ? ~va : va);
        AB::store(acc + AB::STRIDE, j, (idx & N) ? ~vb : vb);
      }
    } else {  // continue a gate: wait for its previous item, then fetch the accumulator (L2 only: another SM wrote it)
      if (tau == 0)
        while (ld_acquire_gpu(A.progress + g) < chunk) __nanosleep(200);
      __syncthreads();
      const uint4* src = reinterpret_cast<const uint4*>(A.scratch + g * (2 * N));
      for (int q = tau; q < 2 * N / 4; q += T) {
        const uint4 v = __ldcg(src + q);
        uint32_t* P = acc + (4 * q >= N ? AB::STRIDE - N : 0);
        AB::store(P, 4 * q + 0, v.x); AB::store(P, 4 * q + 1, v.y); AB::store(P, 4 * q + 2, v.z); AB::store(P, 4 * q + 3, v.w);
      }
    }
    __syncthreads();

    for (int i = i0; i < i1; i++) {
      const int at = abar[i - i0];
      if (at == 0) continue;  // X^0: ct1 - ct0 = 0, digits are all zero, the step is an exact no-op
      if constexpr (TEX)
        cmux_rotate_step<LOGN, L, BGBIT, SMALL>(acc, fft, KeyTex{A.bsk_tex, (int)(row_stride * i)}, at, A.offset, A.tw0);
      else
        cmux_rotate_step<LOGN, L, BGBIT, SMALL>(acc, fft, KeyLdgFor<LOGN>{A.bsk + row_stride * i}, at, A.offset, A.tw0);
      __syncthreads();
    }

    if (chunk + 1 < A.nchunks) {  // hand the accumulator to whoever takes the gate's next item
      uint4* dst = reinterpret_cast<uint4*>(A.scratch + g * (2 * N));
      for (int q = tau; q < 2 * N / 4; q += T) {
        const uint32_t* P = acc + (4 * q >= N ? AB::STRIDE - N : 0) + 4 * q;
        __stcg(dst + q, make_uint4(P[0], P[1], P[2], P[3]));
      }
      __threadfence();
      __syncthreads();
      if (tau == 0) st_release_gpu(A.progress + g, chunk + 1);
    } else {
      br_write_output<N>(A, g, acc, acc + AB::STRIDE, tau, T);
    }
    __syncthreads();  // acc / abar / s_item are rewritten by the next item
  }

  // last block out re-arms the control words for the next launch
  __syncthreads();
  if (tau == 0) {
    __threadfence();
    s_item = (atomicAdd(&A.ctl[1], 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_item) {
    if (A.nchunks > 1)
      for (long long q = tau; q < A.count; q += T) A.progress[q] = 0;
    if (tau == 0) { A.ctl[0] = 0u; A.ctl[1] = 0u; }
  }
}

#if TFHE_EXPERIMENTAL
// =============================================================================================
// TMA-staged variant: the bootstrapping-key row-set of the current (step, digit) is brought into shared memory by
// one bulk asynchronous copy (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier) issued by one thread a
// transform ahead of its use, so the MAC reads the key with short-latency LDS and the L2 latency is hidden
// without spending registers on prefetch.  One 2*M*16-byte buffer (16 KiB at N=1024) is recycled per digit:
// the copy for digit r is issued right after the first block barrier that follows the MAC of digit r-1.
// =============================================================================================
template <int LOGN>
constexpr size_t br_staged_smem_bytes(int n) {
  return (size_t)8 * (1 << LOGN) /*acc*/ + (size_t)(1 << (LOGN - 1)) * 16 /*exchange (single)*/ +
         (size_t)2 * 2 * (1 << (LOGN - 1)) * 16 /*2 key-row buffers (A,B spectra)*/ +
         (size_t)(((n + 1) * 4 + 15) / 16 * 16) /*abar+steps*/ + 32 /*mbarriers*/;
}

template <int LOGN, int L, int BGBIT, bool SMALL, int MINB>
__global__ void __launch_bounds__((1 << (LOGN - 4)), MINB) blind_rotate_staged_kernel(const BrArgs A) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  constexpr uint32_t ROW_BYTES = 2u * M * 16u;           // one digit: A spectrum then B spectrum
  constexpr uint32_t MASK = (BGBIT == 32) ? 0xFFFFFFFFu : ((1u << BGBIT) - 1u);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);                         // [2][N]
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);                    // [M]
  double2* kbuf = reinterpret_cast<double2*>(smem_raw + 8 * N + 16 * M);         // [2 buffers][2][8][T]
  unsigned short* abar = reinterpret_cast<unsigned short*>(smem_raw + 8 * N + 16 * M + 64 * M);
  const int tau = threadIdx.x;
  const long long g = blockIdx.x;
  const int n = A.n;
  unsigned short* steps = abar + n;                                              // indices of the non-trivial steps
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + 8 * N + 16 * M + 64 * M + (((n + 1) * 4 + 15) / 16 * 16));
  uint64_t* full = mbar;        // [2]: key-row buffer b has landed
  uint64_t* rd_bar = mbar + 2;  // exchange-buffer reads done
  __shared__ int s_nsteps;
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
  if (tau == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(rd_bar, T);
  }
  Fft<LOGN - 1, true> fft;
  fft.init(ex, A.tw_tab, tau);
  fft.init_single(rd_bar);
  __syncthreads();
  mbar_arrive(rd_bar);  // completes phase 0: "no reads outstanding" before the first exchange
  if (tau == 0) {  // X^0 steps are exact no-ops (digits all zero): drop them from the schedule
    int c = 0;
    for (int i = 0; i < n; i++)
      if (abar[i] != 0) steps[c++] = (unsigned short)i;
    s_nsteps = c;
  }
  __syncthreads();
  const int nsteps = s_nsteps;
  const int njobs = nsteps * 2 * L;                 // job q = (schedule entry q / 2L, digit q % 2L), buffer q & 1
  const size_t row_stride = (size_t)2 * L * 2 * M;  // double2 per step
  const char* bsk_bytes = reinterpret_cast<const char*>(A.bsk);
  auto issue = [&](int q) {  // thread 0 only
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
    TFHE_UNROLL(TFHE_BR_UNROLL_POLY)
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
      TFHE_UNROLL(TFHE_BR_UNROLL_LVL)
      for (int lvl = 0; lvl < L; lvl++, q++) {
        double2 x[8];
        constexpr double BIAS = 4503599627370496.0 + (double)(1u << (BGBIT - 1));
        const int sh = 32 - (lvl + 1) * BGBIT;
#pragma unroll
        for (int a = 0; a < 8; a++) {
          x[a].x = digit_scaled<BGBIT>(dre[a], sh);
          x[a].y = digit_scaled<BGBIT>(dim[a], sh);
        }
        // The block barrier inside the first exchange proves every thread has finished the MAC of job q-1, whose
        // buffer ((q+1) & 1) is therefore free: stage job q+1 into it, one whole transform ahead of its use.
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
}

#endif  // TFHE_EXPERIMENTAL

// ---------------------------------------------------------------------------------------------
// Stand-alone polynomial transforms in the REFERENCE's FourierPoly layout (groups of 4 real + 4 imaginary parts,
// poly/poly.go:54-62), one block per polynomial — rows a11, a13, a22 of SURVEY.md section 8 at API granularity:
//   mode 0  Evaluator.ToFourierPolyAssign   poly/fourier_transform.go:18-21   u32[N]  -> f64[N]
//   mode 1  Evaluator.ToPolyAssign          poly/fourier_transform.go:31-44   f64[N]  -> u32[N]  (divide by N/2, mod 2^32)
//   mode 2  Evaluator.MulPolyAssign         poly/poly_mul.go:12-22            u32[N] x u32[N] -> u32[N]
// ---------------------------------------------------------------------------------------------
struct PolyArgs {
  const void* in0;
  const void* in1;
  void* out;
  const double2* tw_tab;
  int mode;
  Tw4 tw0;
};

template <int LOGN>
__global__ void __launch_bounds__((1 << (LOGN - 4)), 2) poly_kernel(const PolyArgs A) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* ex = reinterpret_cast<double2*>(smem_raw);  // [2][EXW][M]
  const int tau = threadIdx.x;
  const size_t g = blockIdx.x;
  Fft<LOGN - 1, false> fft;
  fft.init(ex, A.tw_tab, tau);
  __syncthreads();
  auto load_folded = [&](const uint32_t* p, double2 (&x)[8]) {  // fourier_transform.go:64-85
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      x[a].x = (double)(int32_t)p[j];
      x[a].y = (double)(int32_t)p[j + M];
    }
  };
  auto store_poly = [&](const double2 (&x)[8], uint32_t* p) {  // fourier_transform.go:88-125
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = tau + T * a;
      p[j] = to_torus<false>(x[a].x);
      p[j + M] = to_torus<false>(x[a].y);
    }
  };
  const double inv_m = 1.0 / (double)M;
  if (A.mode == 0) {
    double2 x[8];
    load_folded(reinterpret_cast<const uint32_t*>(A.in0) + g * N, x);
    fft.forward(x, A.tw0);
    double* o = reinterpret_cast<double*>(A.out) + g * N;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int k = 8 * tau + e;
      o[(k >> 2) * 8 + (k & 3)] = x[e].x;
      o[(k >> 2) * 8 + 4 + (k & 3)] = x[e].y;
    }
  } else if (A.mode == 1) {
    double2 x[8];
    const double* in = reinterpret_cast<const double*>(A.in0) + g * N;
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int k = 8 * tau + e;
      x[e].x = in[(k >> 2) * 8 + (k & 3)] * inv_m;
      x[e].y = in[(k >> 2) * 8 + 4 + (k & 3)] * inv_m;
    }
    fft.inverse(x, A.tw0);
    store_poly(x, reinterpret_cast<uint32_t*>(A.out) + g * N);
  } else {
    double2 x[8], y[8];
    load_folded(reinterpret_cast<const uint32_t*>(A.in0) + g * N, x);
    fft.forward(x, A.tw0);
    load_folded(reinterpret_cast<const uint32_t*>(A.in1) + g * N, y);
    fft.forward(y, A.tw0);
#pragma unroll
    for (int e = 0; e < 8; e++) {  // poly/fourier_ops.go:138-161, scaled by 1/M for the inverse
      const double re = (x[e].x * y[e].x - x[e].y * y[e].y) * inv_m;
      const double im = (x[e].x * y[e].y + x[e].y * y[e].x) * inv_m;
      x[e] = make_double2(re, im);
    }
    fft.inverse(x, A.tw0);
    store_poly(x, reinterpret_cast<uint32_t*>(A.out) + g * N);
  }
}

// Single CMUX / external product on global-memory TRLWEs (parity-test granularity; rows a9, a14).
template <int LOGN, int L, int BGBIT, bool SMALL, int MINB>
__global__ void __launch_bounds__((1 << (LOGN - 4)), MINB) cmux_kernel(const CmuxArgs A) {
  constexpr int N = 1 << LOGN, M = N / 2, T = M / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw);       // holds ct0, becomes the result
  double2* ex = reinterpret_cast<double2*>(smem_raw + 8 * N);
  uint32_t* c1 = reinterpret_cast<uint32_t*>(smem_raw + 8 * N + 16 * br_nbuf(LOGN) * TFHE_BR_EXW * M);  // [2][N]
  const int tau = threadIdx.x;
  const long long g = blockIdx.x;
  for (int j = tau; j < 2 * N; j += T) {
    acc[j] = A.ct0 ? A.ct0[g * 2 * N + j] : 0u;
    c1[j] = A.ct1[g * 2 * N + j];
  }
  Fft<LOGN - 1> fft;
  fft.init(ex, A.tw_tab, tau);
  __syncthreads();
  // Same data path as cmux_rotate_step but the second operand comes from c1 instead of a rotation.
  double2 accA[8], accB[8];
#pragma unroll
  for (int e = 0; e < 8; e++) { accA[e] = make_double2(0.0, 0.0); accB[e] = make_double2(0.0, 0.0); }
#pragma unroll
  for (int poly = 0; poly < 2; poly++) {
    uint32_t dre[8], dim[8];
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int j = poly * N + tau + T * a;
      dre[a] = c1[j] - acc[j] + A.offset;
      dim[a] = c1[j + M] - acc[j + M] + A.offset;
    }
#pragma unroll
    for (int lvl = 0; lvl < L; lvl++) {
      double2 x[8];
      const int sh = 32 - (lvl + 1) * BGBIT;
#pragma unroll
      for (int a = 0; a < 8; a++) {
        x[a].x = digit_scaled<BGBIT>(dre[a], sh);
        x[a].y = digit_scaled<BGBIT>(dim[a], sh);
      }
      fft.forward(x, A.tw0);
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const double2 ka = __ldg(A.bsk_row + key_pos<T>(poly * L + lvl, 0, e, tau));
        const double2 kb = __ldg(A.bsk_row + key_pos<T>(poly * L + lvl, 1, e, tau));
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
  uint32_t* o = A.out + g * 2 * N;
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
    o[j] = acc[j] + to_torus<SMALL>(accA[a].x);
    o[j + M] = acc[j + M] + to_torus<SMALL>(accA[a].y);
  }
  fft.inverse(accB, A.tw0);
#pragma unroll
  for (int a = 0; a < 8; a++) {
    const int j = tau + T * a;
    o[N + j] = acc[N + j] + to_torus<SMALL>(accB[a].x);
    o[N + j + M] = acc[N + j + M] + to_torus<SMALL>(accB[a].y);
  }
}

}  // namespace tfhe
