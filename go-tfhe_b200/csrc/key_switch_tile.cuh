// Identity key switch for the large-base sets (Uint2-5: base = 16..64), batches of a few hundred ciphertexts and more.
//
//   out_c = (0,...,0,b_c) - sum_{i<N, j<t} KSK[base*(i*t+j) + k_c(i,j)]        (trgsw/keyswitch.go:10-37)
//
// The row gather (key_switch_kernel) reads every selected row once PER CIPHERTEXT out of L2: 2048 Uint5 ciphertexts pull
// 52 GB through the L2->SM fabric (3.5 ms, profiles/r02_ks_gather_uint5_ncu_full_metrics.csv).  A dense one-hot
// contraction on the tensor cores (the basebit = 2 path) would be base-1 = 63 times the work.  Here a block owns a TILE
// of 256 ciphertexts x 64 output words and walks the (i, j) pairs; for each pair the `base` candidate rows' 64-word
// column slices are staged ONCE in shared memory (one 32 KiB TMA box of 128 key rows = 128 / base pairs per stage, 3 stages in flight on mbarriers, refilled by whichever warp finishes a buffer last: no block barrier in the loop) and each of the 256
// ciphertexts adds the slice its digit selects: every key byte leaves L2 once per 256 ciphertexts instead of once per
// ciphertext (13.7 GB instead of 52 GB), the selection itself becomes a shared-memory read.  Sums are u32 and commute:
// bit-identical to the gather and to the oracle.  K = N*t pairs are split over several blocks per tile (red.global.add
// into the pre-initialised output) so that the grid fills whole waves of SMs.
//
// Digits come from a pre-pass (ks_digits_kernel): D[pair][ciphertext] u8, ciphertexts padded to the tile with digit 0,
// whose rows are all-zero in the device key (ksk_repack_kernel).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "blind_rotate.cuh"    // mbarrier helpers
#include "key_switch_mma.cuh"  // tma_load_2d
#include "lwe_kernels.cuh"

namespace tfhe {

constexpr int KST_CT = 256;      // ciphertexts per tile
constexpr int KST_COLS = 64;     // output words per tile (16 uint4)
constexpr int KST_THREADS = 512; // 16 column quads x 32 ciphertext groups of 8
constexpr int KST_STAGES = 3;
constexpr int KST_STAGE_ROWS = 128;  // key rows per stage = 128 / base (i, j) pairs: 32 KiB, one TMA box
__host__ __device__ constexpr size_t kst_smem_bytes(int) { return (size_t)KST_STAGES * KST_STAGE_ROWS * KST_COLS * 4; }

// lwe_in [count][N+1] -> digits D[(i*t+j)][cpad] and out rows initialised to (0,...,0,b).
// grid (cpad / 256, N / 8), 256 threads: thread = ciphertext, blockIdx.y = group of 8 mask words (one 32-byte sector).
__global__ void __launch_bounds__(256) ks_digits_kernel(const uint32_t* __restrict__ lwe_in, uint8_t* __restrict__ D,
                                                        uint32_t* __restrict__ out, long long count, long long cpad, int N, int n,
                                                        int basebit, int t, const GateDesc* __restrict__ out_gates,
                                                        long long instances, long long g_base) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  const bool live = c < count;
  const uint32_t prec = 1u << (32 - (1 + basebit * t));
  const uint32_t mask = (1u << basebit) - 1u;
  const uint32_t* src = lwe_in + (size_t)(live ? c : 0) * (N + 1);
  const int i0 = blockIdx.y * 8;
#pragma unroll
  for (int u = 0; u < 8; u++) {
    const int i = i0 + u;
    const uint32_t abar = (live ? src[i] : 0u - prec) + prec;  // padding ciphertexts: every digit 0
    for (int j = 0; j < t; j++)
      D[(size_t)(i * t + j) * cpad + c] = (uint8_t)((abar >> (32 - (j + 1) * basebit)) & mask);
  }
  if (live) {  // this block's slice of the output row
    const long long g = g_base + c;  // job index in the whole batch (lwe_in and D are chunk-local, out is not)
    const size_t orow = out_gates ? (size_t)out_gates[g / instances].out * instances + (size_t)(g % instances) : (size_t)g;
    uint32_t* o = out + orow * (n + 1);
    const int per = (n + 1 + gridDim.y - 1) / gridDim.y;
    const int w0 = blockIdx.y * per, w1 = min(n + 1, w0 + per);
    for (int w = w0; w < w1; w++) o[w] = (w == n) ? src[N] : 0u;
  }
}

// grid = col_tiles * ksplit * ct_tiles; blockIdx.x = (col_tile * ksplit + ks) * ct_tiles + ct_tile: the blocks that stream
// the same key slice (all ciphertext tiles of one column tile and pair range) are neighbours in launch order, run at the
// same time and share each slab out of L2 — the key leaves HBM about once per launch.
template <int BASE>
__global__ void __launch_bounds__(KST_THREADS, 2) ks_tile_kernel(const __grid_constant__ CUtensorMap key_map, const uint8_t* __restrict__ D,
                                                                 uint32_t* __restrict__ out, int K, long long count, long long cpad,
                                                                 int n, int col_tiles, int ksplit, int ct_tiles,
                                                                 const GateDesc* __restrict__ out_gates, long long instances,
                                                                 long long g_base) {
  extern __shared__ __align__(128) uint4 slab[];  // [KST_STAGES][KST_STAGE_ROWS][16]
  __shared__ __align__(8) uint64_t full_bar[KST_STAGES];
  __shared__ int done[KST_STAGES];                 // warps that have finished reading each buffer
  constexpr int PP = KST_STAGE_ROWS / BASE;        // (i, j) pairs per stage
  constexpr uint32_t STAGE_BYTES = KST_STAGE_ROWS * 16 * 16;
  constexpr int WARPS = KST_THREADS / 32;
  const int tid = threadIdx.x;
  const int ct_tile = blockIdx.x % ct_tiles;
  const int rest = blockIdx.x / ct_tiles;
  const int ks = rest % ksplit, col_tile = rest / ksplit;
  // K counts stages (groups of PP pairs) here
  const int s_lo = (int)(((long long)K * ks) / ksplit), s_hi = (int)(((long long)K * (ks + 1)) / ksplit);
  const int nst = s_hi - s_lo;
  const int colq = tid & 15, ctg = tid >> 4;
  const long long c0 = (long long)ct_tile * KST_CT + ctg * 8;
  const uint32_t smem0 = smem_u32(slab);
  if (tid == 0) {
#pragma unroll
    for (int p = 0; p < KST_STAGES; p++) { mbar_init(&full_bar[p], 1); done[p] = 0; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // one TMA box per stage: 128 rows x 64 words of the key ([rows][stride] u32), columns past the row end arrive as zeros
  auto fill = [&](int s, int buf) {
    mbar_arrive_expect_tx(&full_bar[buf], STAGE_BYTES);
    tma_load_2d(smem0 + buf * STAGE_BYTES, &key_map, col_tile * KST_COLS, (s_lo + s) * KST_STAGE_ROWS, smem_u32(&full_bar[buf]));
  };
  if (tid == 0)
    for (int p = 0; p < KST_STAGES && p < nst; p++) fill(p, p);
  const uint2* dig = reinterpret_cast<const uint2*>(D + (size_t)s_lo * PP * cpad + c0);
  const size_t dig_step = (size_t)cpad / 8;  // uint2 per pair
  uint4 acc[8];
#pragma unroll
  for (int u = 0; u < 8; u++) acc[u] = make_uint4(0u, 0u, 0u, 0u);
  uint2 dnext = nst > 0 ? __ldg(dig) : make_uint2(0u, 0u);
  int buf = 0;
  uint32_t phase = 0;
  // No block-wide barrier in the loop: a warp waits only for its stage's data; the LAST warp to finish reading a buffer
  // (counted in shared memory) refills it with the stage KST_STAGES ahead, so fast warps run up to two stages ahead of slow ones.
  for (int s = 0; s < nst; s++) {
    mbar_wait(&full_bar[buf], phase);  // stage s has landed
#pragma unroll 1
    for (int pp = 0; pp < PP; pp++) {
      const uint2 d = dnext;
      dig += dig_step;
      if (pp + 1 < PP || s + 1 < nst) dnext = __ldg(dig);
      const unsigned char* rd = reinterpret_cast<const unsigned char*>(slab) + colq * 16 + buf * STAGE_BYTES + pp * (BASE * 256);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        // byte u of the digit word, times the 256-byte row pitch: one byte permute puts it in bits 8..15
        const uint32_t off = __byte_perm(u < 4 ? d.x : d.y, 0u, 0x4404u | ((uint32_t)(u & 3) << 4));
        const uint4 v = *reinterpret_cast<const uint4*>(rd + off);
        acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
      }
    }
    __syncwarp();  // every lane's reads of this buffer have been issued and consumed
    if ((tid & 31) == 0 && s + KST_STAGES < nst) {
      if (atomicAdd(&done[buf], 1) == WARPS - 1) {
        done[buf] = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fill(s + KST_STAGES, buf);
      }
    }
    if (++buf == KST_STAGES) { buf = 0; phase ^= 1u; }
  }
  const int cw = col_tile * KST_COLS + colq * 4;
#pragma unroll
  for (int u = 0; u < 8; u++) {
    const long long c = c0 + u;
    if (c >= count) continue;
    const long long g = g_base + c;
    const size_t orow = out_gates ? (size_t)out_gates[g / instances].out * instances + (size_t)(g % instances) : (size_t)g;
    uint32_t* o = out + orow * (n + 1) + cw;
    if (cw + 0 <= n) atomicAdd(o + 0, 0u - acc[u].x);
    if (cw + 1 <= n) atomicAdd(o + 1, 0u - acc[u].y);
    if (cw + 2 <= n) atomicAdd(o + 2, 0u - acc[u].z);
    if (cw + 3 <= n) atomicAdd(o + 3, 0u - acc[u].w);
  }
}

}  // namespace tfhe
