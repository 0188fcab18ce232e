// key_switch_mma.cuh — identity key switch of a whole batch as ONE exact integer contraction on the 5th-generation
// tensor cores (tcgen05.mma kind::i8, accumulators in tensor memory, operands staged by TMA).
//
// Reference: trgsw/keyswitch.go:10-37 (= trgsw/trgsw.go:285-311)
//   out = (0,...,0,b) - sum_{i<N, j<t, k_ij != 0} KSK[(base*t*i) + (base*j) + k_ij]        (all n+1 words, mod 2^32)
// For a batch this is   out[g][w] = b_g [w == n]  -  sum_K  S[g][K] * KSK[K][w],   K = (i, j, k >= 1),
// with S the 0/1 selection matrix (one 1 per (i, j) whose digit is non-zero).  Splitting every u32 key word into its
// four bytes turns the sum into a u8 x u8 -> s32 matrix product (at most N*t terms of <= 255: no overflow), and
//   sum_K S*KSK[K][w] = sum_p 2^(8p) * D[g][4w + p]   (mod 2^32)
// so the result is bit-identical to the row-gather kernel (key_switch_kernel) — additions mod 2^32 commute.
//
// Why: the gather kernel moves N*t*(1 - 1/base) rows (19.4 MB @128-bit) per ciphertext from L2 and sits on the
// L2->SM bandwidth (18 TB/s, 4.4 ms per 4096).  Here every key byte is loaded once per 128-ciphertext tile and
// reused out of shared memory by the tensor cores; the dense product costs base-1 = 3x the additions, which is
// why this path is only taken for basebit = 2 (the 80/110/128-bit gate sets).
//
// Layouts (all K-major, 128-byte swizzle, K = N*t*(base-1), kidx = (i*t + j)*(base-1) + (k-1)):
//   A  (selection)  [count][K]        u8   written per call by ks_onehot_kernel
//   Bt (key bytes)  [4*stride][K]     u8   written once at key load by ksk_bytes_repack_kernel; row 4w+p = byte p of word w
//   D  tile 128 ciphertexts x BN byte-columns, s32, in TMEM (lane = ciphertext, column = byte-column)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "blind_rotate.cuh"   // mbarrier helpers
#include "lwe_kernels.cuh"    // GateDesc

namespace tfhe {

constexpr int KSM_BM = 128;        // ciphertexts per tile (TMEM lanes)
constexpr int KSM_BN = 256;        // byte-columns per tile (TMEM columns)
constexpr int KSM_BK = 128;        // K bytes per pipeline stage (= one 128-byte swizzle row)
constexpr int KSM_STAGES = 4;
constexpr int KSM_A_BYTES = KSM_BM * KSM_BK;   // 16 KiB
constexpr int KSM_B_BYTES = KSM_BN * KSM_BK;   // 32 KiB
constexpr int KSM_THREADS = 192;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr size_t KSM_SMEM = (size_t)KSM_STAGES * (KSM_A_BYTES + KSM_B_BYTES) + 1024 /*alignment slack*/;

// ---- one-time: key rows -> byte planes, K-major --------------------------------------------------------------
// src: [N*t*base][stride] u32 (reference row order, k = 0 rows included).  dst: [4*stride][K] u8.
// One block transposes a tile of 64 kidx x 64 words through shared memory.
__global__ void __launch_bounds__(256) ksk_bytes_repack_kernel(const uint32_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                               int stride, int basebit, long long K) {
  __shared__ uint32_t tile[64][65];
  const int bm1 = (1 << basebit) - 1;
  const long long k0 = (long long)blockIdx.x * 64;
  const int w0 = blockIdx.y * 64;
  for (int q = threadIdx.x; q < 64 * 64; q += 256) {
    const int kk = q >> 6, w = q & 63;
    const long long kidx = k0 + kk;
    uint32_t v = 0;
    if (kidx < K && w0 + w < stride) {
      const long long ij = kidx / bm1;
      const int k = (int)(kidx - ij * bm1) + 1;
      v = src[(size_t)((ij << basebit) + k) * stride + w0 + w];
    }
    tile[kk][w] = v;
  }
  __syncthreads();
  // 256 byte-columns x 64 kidx bytes: each thread writes 16 consecutive kidx bytes of one column per pass
  for (int q = threadIdx.x; q < 256 * 4; q += 256) {
    const int col = q >> 2, part = q & 3;
    const int w = col >> 2, p = col & 3;
    if (w0 + w >= stride) continue;
    uint32_t o[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      uint32_t pk = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) pk |= ((tile[part * 16 + u * 4 + b][w] >> (8 * p)) & 0xFFu) << (8 * b);
      o[u] = pk;
    }
    const long long kk0 = k0 + part * 16;
    if (kk0 + 16 <= K)
      *reinterpret_cast<uint4*>(dst + (size_t)((w0 + w) * 4 + p) * K + kk0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- per call: selection matrix + output initialisation ----------------------------------------------------
// grid = count, block = 256, dynamic shared memory = K bytes.
__global__ void __launch_bounds__(256) ks_onehot_kernel(const uint32_t* __restrict__ lwe_in, uint8_t* __restrict__ A,
                                                        uint32_t* __restrict__ out, int N, int n, int basebit, int t,
                                                        int K, const GateDesc* __restrict__ out_gates,
                                                        long long instances, long long g_base) {
  extern __shared__ __align__(16) uint8_t sel[];
  const long long g = blockIdx.x;  // index inside this chunk; g_base + g is the job index of the whole batch
  const uint32_t* src = lwe_in + (size_t)g * (N + 1);
  uint4* sel4 = reinterpret_cast<uint4*>(sel);
  for (int q = threadIdx.x; q < K / 16; q += blockDim.x) sel4[q] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const uint32_t prec = 1u << (32 - (1 + basebit * t));
  const uint32_t mask = (1u << basebit) - 1u;
  const int bm1 = (1 << basebit) - 1;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const uint32_t abar = src[i] + prec;
    for (int j = 0; j < t; j++) {
      const uint32_t k = (abar >> (32 - (j + 1) * basebit)) & mask;
      if (k != 0) sel[(i * t + j) * bm1 + (int)k - 1] = 1;
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(A + (size_t)g * K);
  for (int q = threadIdx.x; q < K / 16; q += blockDim.x) dst[q] = sel4[q];
  const long long job = g_base + g;
  const size_t orow = out_gates ? (size_t)out_gates[job / instances].out * instances + (size_t)(job % instances) : (size_t)job;
  uint32_t* o = out + orow * (n + 1);
  const uint32_t bterm = src[N];
  for (int w = threadIdx.x; w <= n; w += blockDim.x) o[w] = (w == n) ? bterm : 0u;
}

// ---- tcgen05 / TMA wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (sm_100 descriptor version 1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

struct KsMmaArgs {
  uint32_t* out;                 // [.][n+1], pre-initialised to (0,...,0,b) by ks_onehot_kernel
  const GateDesc* out_gates;     // circuits: job -> wire row (see key_switch_kernel)
  long long instances;
  long long g_base;              // job index of this chunk's first ciphertext
  int count;                     // ciphertexts in this chunk
  int n;
  int kblocks;                   // K / 128
  int ksplit;                    // K range split across this many work items (results combine by atomic add)
  int tiles_m, tiles_n;
};

// grid = tiles_m * tiles_n * ksplit, block = 192, one CTA per SM (4 x 48 KiB stages).
__global__ void __launch_bounds__(KSM_THREADS, 1) ks_mma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapB, const KsMmaArgs a) {
  extern __shared__ uint8_t ksm_raw[];
  __shared__ __align__(8) uint64_t full_bar[KSM_STAGES], empty_bar[KSM_STAGES], tmem_full_bar;
  __shared__ uint32_t s_tmem_base;
  const uint32_t smem_base = (smem_u32(ksm_raw) + 1023u) & ~1023u;  // 128-byte swizzle atoms need 1024-byte alignment
  const uint32_t smem_a = smem_base, smem_b = smem_base + KSM_STAGES * KSM_A_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tiles = a.tiles_m * a.tiles_n;
  const int ks = blockIdx.x / tiles;
  const int rem = blockIdx.x - ks * tiles;
  const int mt = rem / a.tiles_n, nt = rem - mt * a.tiles_n;
  const int kb0 = (int)(((long long)a.kblocks * ks) / a.ksplit), kb1 = (int)(((long long)a.kblocks * (ks + 1)) / a.ksplit);
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < KSM_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tmem_full_bar, 1);
  }
  if (warp == 2) {  // one warp owns the TMEM allocation (256 columns: the whole 128 x 256 s32 accumulator tile)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int it = 0; it < nkb; it++) {
        const int s = it % KSM_STAGES;
        const uint32_t ph = (uint32_t)(it / KSM_STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);  // slot released by the MMA that read it (passes at once the first time round)
        mbar_arrive_expect_tx(&full_bar[s], KSM_A_BYTES + KSM_B_BYTES);
        const int kbyte = (kb0 + it) * KSM_BK;
        tma_load_2d(smem_a + s * KSM_A_BYTES, &mapA, kbyte, mt * KSM_BM, smem_u32(&full_bar[s]));
        tma_load_2d(smem_b + s * KSM_B_BYTES, &mapB, kbyte, nt * KSM_BN, smem_u32(&full_bar[s]));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer: one thread drives the tensor core for the whole CTA =====
      // instruction descriptor (kind::i8): D = s32, A = B = unsigned 8-bit, both K-major, N = 256, M = 128
      constexpr uint32_t IDESC = (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(KSM_BN >> 3) << 17) | ((uint32_t)(KSM_BM >> 4) << 24);
      for (int it = 0; it < nkb; it++) {
        const int s = it % KSM_STAGES;
        const uint32_t ph = (uint32_t)(it / KSM_STAGES) & 1u;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_a + s * KSM_A_BYTES, sb = smem_b + s * KSM_B_BYTES;
#pragma unroll
        for (int k = 0; k < KSM_BK / 32; k++)  // UMMA_K = 32 bytes for 8-bit operands: advance inside the swizzle row
          umma_i8(tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), IDESC, (uint32_t)((it | k) != 0));
        umma_commit(&empty_bar[s]);  // arrives when the MMAs above have finished reading the slot
      }
      umma_commit(&tmem_full_bar);
    }
  } else {  // ===== epilogue: TMEM -> registers -> recombine the four byte planes -> atomic subtract =====
    mbar_wait(&tmem_full_bar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                       // a warp may only touch TMEM lanes [32 q, 32 q + 32)
    const int row = mt * KSM_BM + q * 32 + lane;  // ciphertext of this thread
    const bool live = row < a.count;
    const long long job = a.g_base + row;
    const size_t orow = (live && a.out_gates) ? (size_t)a.out_gates[job / a.instances].out * a.instances + (size_t)(job % a.instances)
                                              : (size_t)job;
    uint32_t* o = a.out + orow * (a.n + 1);
    const int w_tile = nt * (KSM_BN / 4);
#pragma unroll 1
    for (int c = 0; c < KSM_BN / 32; c++) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (live && nkb > 0) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int w = w_tile + c * 8 + u;
          const uint32_t sum = v[4 * u] + (v[4 * u + 1] << 8) + (v[4 * u + 2] << 16) + (v[4 * u + 3] << 24);
          if (w <= a.n && sum != 0u) atomicAdd(o + w, 0u - sum);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace tfhe
