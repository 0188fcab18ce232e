// tfhe_b200.cu — context, key upload and the C ABI of include/tfhe_b200.h.
// Host side of the engine: owns device memory, launches the kernels of blind_rotate.cuh and
// lwe_kernels.cuh.  No CPU fallback anywhere: if CUDA is unavailable every call fails.
#include "../../include/tfhe_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

// TFHE_EXPERIMENTAL: also builds the measured-slower blind-rotate variants of round 1 (TMA-staged / texture key fetch,
// TMEM accumulators, TMEM exchange, warp-per-gate, gates-per-block, per-digit and cluster latency kernels; see
// profiles/r01_experiments.md).  The default library holds only what the engine selects by itself: the throughput
// kernel, the two latency kernels (lat, latp), both key-switch kernels and key generation.
#ifndef TFHE_EXPERIMENTAL
#define TFHE_EXPERIMENTAL 0
#endif
#include "blind_rotate.cuh"
#if TFHE_EXPERIMENTAL
#include "blind_rotate_w16.cuh"
#include "blind_rotate_tx.cuh"
#include "blind_rotate_tms.cuh"
#include "blind_rotate_mg.cuh"
#endif
#include "blind_rotate_lat.cuh"
#include "lwe_kernels.cuh"
#include "key_switch_mma.cuh"
#include "key_switch_tile.cuh"
#include "keygen.cuh"
#include "fp64_probe.cuh"

using namespace tfhe;

namespace {

std::string g_create_error;
std::mutex g_create_mu;

// bumped whenever a device buffer moves: captured CUDA graphs hold raw pointers and are re-captured after any move
static std::atomic<uint64_t> g_alloc_generation{1};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    g_alloc_generation.fetch_add(1, std::memory_order_relaxed);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) { cudaFree(p); g_alloc_generation.fetch_add(1, std::memory_order_relaxed); } p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// staging slot of the pipelined host-buffer API: device copies of one chunk's inputs / outputs, the events that order
// the three streams, and a pinned host buffer for the index lists of a mixed gate batch
struct Slot {
  DevBuf a, b, c, luts, out;
  cudaEvent_t in_done = nullptr, comp_done = nullptr, out_done = nullptr, idx_done = nullptr;
  int* h_idx = nullptr;
  size_t h_idx_cap = 0;
  bool idx_pending = false;
};

}  // namespace

struct tfhe_ctx {
  tfhe_params P{};
  // a context created by tfhe_ctx_create_multi owns one single-device context per GPU and no device state of its own
  std::vector<tfhe_ctx*> kids;
  int device = 0;
  int logN = 0;
  int variant = -1;          // index into the kernel instantiation table
  uint32_t offset = 0;
  bool key_loaded = false, has_ksk = false;
  double2* d_bsk = nullptr;  // [n][2L][2][8][T]
  uint32_t* d_ksk = nullptr; // [N*t*base][ksk_stride]
  uint32_t* d_testvec = nullptr;
  double2* d_tw = nullptr;
  int ksk_stride = 0;
  // tensor-core key switch (key_switch_mma.cuh): byte planes of the key, K-major, and its TMA descriptor
  uint8_t* d_ksk_bytes = nullptr;  // [4*ksk_stride][ks_K]
  long long ks_K = 0;              // N * t * (base - 1); 0 = path unavailable for this parameter set
  CUtensorMap ks_mapB{};
  CUtensorMap ks_tile_map{};      // the repacked key-switching key as [rows][stride] u32, box = base rows x 64 words (key_switch_tile.cuh)
  bool ks_tile_ok = false;
  int ks_variant = 0;              // 0 = auto, 1 = row gather (key_switch_kernel), 2 = tensor-core contraction, 3 = shared-memory tiles (ks_tile_kernel)
  int br_tail_min = 6;            // shortest piece of the graded tail of a gate's work items (steps); >= n: no grading
  int mux_mode = 0;                // 0 = the reference's three bootstraps per MUX, 1 = two blind rotations + one key switch
  // proxy re-encryption key (proxyreenc.ProxyReencryptionKey.KeyEncryptions): [n*t*base][stride] rows under the target key
  uint32_t* d_reenc = nullptr;
  int reenc_stride = 0, reenc_basebit = 0, reenc_t = 0;
  DevBuf ks_sel;                   // selection matrix of the current chunk (tensor-core path) / digit matrix (tile path)
  Tw4 tw0{};
  cudaStream_t stream = nullptr;  // used by the host-buffer API (compute)
  cudaStream_t s_in = nullptr, s_out = nullptr;  // copy-in / copy-out streams of the pipelined host-buffer calls
  Slot slot[2];
  int64_t pipe_chunk = 16384;     // ciphertexts per pipeline chunk
  DevBuf prep, lwe1, tmp, prep2, idx_a, idx_b, ops_dev;       // engine scratch
  DevBuf wires, gate_descs;                                   // circuit runner
  // circuit runner, CUDA-graph replay (tfhe_ctx_set_circuit_graph): the level loop of a circuit that was already run once
  // with the same shape is captured on the next call and replayed afterwards
  struct CircuitGraph {
    std::vector<tfhe_gate_desc> gates;
    int32_t n_inputs = 0;
    int64_t instances = 0;
    uint64_t gen = 0;            // g_alloc_generation at capture time (0: seen once, not captured yet)
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  std::vector<CircuitGraph> circuit_graphs;
  int circuit_graph = 0;          // 0 = off (default), 1 = capture and replay
  int64_t circuit_graph_replays = 0;
  DevBuf h2d_a, h2d_b, h2d_c, h2d_luts, d2h_out;              // staging for the host-buffer API
  int64_t launches = 0;
  int sm_count = 0;
  // persistent throughput kernel: resident blocks per SM, control words (ctl[2] + progress[]) and accumulator hand-over
  int br_blocks_per_sm = 1;
  int br_chunk_steps = 0;    // 0 = automatic (launch_blind_rotate), else CMUX steps per work item (>= n: whole gates)
  DevBuf br_ctl, br_scratch;
  std::string err;
  // optional per-stage timing (tfhe_ctx_set_timing): CUDA events on the launching stream
  // how blind rotate reads the key rows: 0 = LDG straight from L2 (default, fastest measured), 1 = TMA-staged through
  // shared memory (cp.async.bulk + mbarrier), 2 = texture fetches.  See profiles/ for the measurements behind the default.
  int br_variant = 0;
  bool br_auto_cl = false;   // variant 0 may also pick the cluster kernel for a handful of gates (set once measured)
  bool br_auto_lat = true;   // variant 0 picks the latency kernel for batches of <= 2 gates per SM (exact parameter sets)
  cudaTextureObject_t bsk_tex = 0;
#if TFHE_EXPERIMENTAL
  // warp-per-gate kernel (N = 1024 only): its own key layout and twiddle tables
  double2* d_bsk16 = nullptr;
  double2* d_tw16 = nullptr;   // [8][16] pass-1 twiddles then [4][32] last-stage twiddles
  Tw8 tw0_16{};
#endif
  bool timing = false;
  struct StageEv { cudaEvent_t e0, e1, e2; };
  std::vector<StageEv> ev_live, ev_free;
};

namespace {

int fail(tfhe_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else { std::lock_guard<std::mutex> l(g_create_mu); g_create_error = buf; }
  return code;
}
#define CK(c, call)                                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) return fail((c), TFHE_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// S(m, i): twiddle of block i at the stage with m blocks — the square root the reference tabulates
// as tw[m-1+i] (poly/poly_evaluator.go:114-143): exp(-2 pi i rev(i)/M) * exp(i pi / (4m)), rev = bit
// reversal over log2(M/2) bits.  Evaluated in long double and rounded once.
double2 twiddle(int M, int m, int i) {
  int bits = 0;
  while ((1 << bits) < M / 2) bits++;
  int r = 0;
  for (int b = 0; b < bits; b++)
    if (i & (1 << b)) r |= 1 << (bits - 1 - b);
  const long double pi = 3.14159265358979323846264338327950288L;
  long double ang = -2.0L * pi * (long double)r / (long double)M + pi / (4.0L * (long double)m);
  return make_double2((double)cosl(ang), (double)sinl(ang));
}
Tw4 block_twiddles(int M, int m, int i) {
  Tw4 t;
  t.s[0] = twiddle(M, m, i);
  t.s[1] = (2 * m <= M / 2) ? twiddle(M, 2 * m, 2 * i) : make_double2(1, 0);
  t.s[2] = (4 * m <= M / 2) ? twiddle(M, 4 * m, 4 * i) : make_double2(1, 0);
  t.s[3] = (4 * m <= M / 2) ? twiddle(M, 4 * m, 4 * i + 2) : make_double2(1, 0);
  return t;
}
template <int LOGM>
void build_twiddles(Tw4& tw0, std::vector<Tw4>& tab) {
  using G = Geo<LOGM>;
  tw0 = block_twiddles(G::M, 1, 0);
  tab.assign(G::tab_len() > 0 ? G::tab_len() : 1, Tw4{});
  for (int k = 1; k < G::NPASS; k++) {
    const int m = (k < G::NFULL) ? (1 << (3 * k)) : G::T;  // partial pass: virtual (m, i) = (M/8, tau)
    for (int i = 0; i < G::blocks(k); i++) tab[G::tab_off(k) + i] = block_twiddles(G::M, m, i);
  }
}

// ---- kernel instantiation table ---------------------------------------------------------------
struct Variant {
  int logN, L, bgbit;
  bool small;
  void (*br)(const BrArgs);          // throughput kernel: block per gate, key rows read straight from L2 (LDG)
  void (*cmux)(const CmuxArgs);
  size_t (*br_smem)(int n);
  void (*br_lat)(const BrArgs);      // latency mode: 4 warps per gate, the two polynomials in parallel (exact sets, N = 1024)
  size_t (*br_lat_smem)(int n);
  void (*br_latp)(const BrArgs);     // latency mode that keeps the reference's accumulation order (L <= 2: the Uint / PBS sets)
  size_t (*br_latp_smem)(int n);
#if TFHE_EXPERIMENTAL
  void (*br_tex)(const BrArgs);      // key rows fetched through the texture pipe
  void (*br_staged)(const BrArgs);   // key rows TMA-staged into shared memory (cp.async.bulk + mbarrier)
  size_t (*br_staged_smem)(int n);
  void (*br_w16)(const BrW16Args);   // warp-per-gate, TMEM accumulators (N = 1024 only, else nullptr)
  void (*br_tm)(const BrArgs);       // block-per-gate, TMEM accumulators (N >= 1024, else nullptr)
  void (*br_tx)(const BrArgs);       // block-per-gate, second transform exchange through TMEM (N = 1024, else nullptr)
  void (*br_txs)(const BrArgs);      // same + key rows TMA-staged through shared memory
  void (*br_tms)(const BrArgs);      // TMEM accumulators + TMA-staged key rows + single exchange buffer: 6 blocks/SM (N = 1024)
  size_t (*br_tms_smem)(int n);
  void (*br_mg)(const BrArgs, long long);  // G gates per block sharing one staged copy of the key rows, TMEM accumulators (N = 1024)
  size_t (*br_mg_smem)(int n);
  void (*br_lat2)(const BrArgs);     // latency mode, one group per digit (2L x N/16 threads): <= 1 gate per SM
  size_t (*br_lat2_smem)(int n);
  void (*br_cl)(const BrArgs);       // one gate on a cluster of 2L blocks (one SM per digit, DSMEM): a handful of gates
  size_t (*br_cl_smem)(int n);
#endif
};
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto latp_kernel() -> void (*)(const BrArgs) {
  if constexpr (L <= 2) return blind_rotate_latp_kernel<LOGN, L, BG, SMALL>;
  else return nullptr;
}
template <int LOGN, int L> size_t br_latp_smem(int n) { return br_latp_smem_bytes<LOGN, (L <= 2 ? L : 1)>(n); }
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto lat_kernel() -> void (*)(const BrArgs) {
  if constexpr (LOGN == 10 && SMALL) return blind_rotate_lat_kernel<LOGN, L, BG, SMALL>;
  else return nullptr;
}
template <int LOGN> size_t br_lat_smem(int n) { return br_lat_smem_bytes<LOGN>(n); }
template <int LOGN> size_t br_smem(int n) { return br_smem_bytes<LOGN>(n); }
#if TFHE_EXPERIMENTAL
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto cl_kernel() -> void (*)(const BrArgs) {
  if constexpr (LOGN == 10 && SMALL && 2 * L <= 8) return blind_rotate_cl_kernel<LOGN, L, BG, SMALL>;
  else return nullptr;
}
template <int LOGN> size_t br_cl_smem(int n) { return br_cl_smem_bytes<LOGN>(n); }
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto lat2_kernel() -> void (*)(const BrArgs) {
  if constexpr (LOGN == 10 && SMALL) return blind_rotate_lat2_kernel<LOGN, L, BG, SMALL>;
  else return nullptr;
}
template <int LOGN, int L> size_t br_lat2_smem(int n) { return br_lat2_smem_bytes<LOGN, L>(n); }
#ifndef TFHE_BR_MG_G
#define TFHE_BR_MG_G 6
#endif
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto mg_kernel() -> void (*)(const BrArgs, long long) {
  if constexpr (LOGN == 10) return blind_rotate_mg_kernel<LOGN, L, BG, SMALL, TFHE_BR_MG_G>;
  else return nullptr;
}
template <int LOGN> size_t br_mg_smem(int n) { return br_mg_smem_bytes<LOGN, TFHE_BR_MG_G>(n); }
#ifndef TFHE_BR_TMS_MINB
#define TFHE_BR_TMS_MINB 6
#endif
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto tms_kernel() -> void (*)(const BrArgs) {
  if constexpr (LOGN == 10) return blind_rotate_tms_kernel<LOGN, L, BG, SMALL, TFHE_BR_TMS_MINB>;
  else return nullptr;
}
template <int LOGN> size_t br_tms_smem(int n) { return br_tms_smem_bytes<LOGN>(n); }
template <int LOGN> size_t br_staged_smem(int n) { return br_staged_smem_bytes<LOGN>(n); }
#ifndef TFHE_BR_W16_MINB
#define TFHE_BR_W16_MINB 2
#endif
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto w16_kernel() -> void (*)(const BrW16Args) {
  if constexpr (LOGN == 10) return blind_rotate_w16_kernel<L, BG, SMALL, TFHE_BR_W16_MINB>;
  else return nullptr;
}
#ifndef TFHE_BR_TX_MINB
#define TFHE_BR_TX_MINB 4
#endif
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto tx_kernel() -> void (*)(const BrArgs) {
  if constexpr (LOGN == 10) return blind_rotate_tx_kernel<L, BG, SMALL, TFHE_BR_TX_MINB>;
  else return nullptr;
}
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto txs_kernel() -> void (*)(const BrArgs) {
  if constexpr (LOGN == 10) return blind_rotate_txs_kernel<L, BG, SMALL, TFHE_BR_TX_MINB>;
  else return nullptr;
}
#ifndef TFHE_BR_TM_MINB_N1024
#define TFHE_BR_TM_MINB_N1024 6
#endif
template <int LOGN, int L, int BG, bool SMALL>
constexpr auto tm_kernel() -> void (*)(const BrArgs) {
  if constexpr (LOGN == 10) return blind_rotate_tm_kernel<LOGN, L, BG, SMALL, TFHE_BR_TM_MINB_N1024>;
  else if constexpr (LOGN == 11) return blind_rotate_tm_kernel<LOGN, L, BG, SMALL, 3>;
  else return nullptr;
}
#ifndef TFHE_BR_STAGED_MINB_N1024
#define TFHE_BR_STAGED_MINB_N1024 4
#endif
#define VARIANT_EXP(LOGN, L, BG, SMALL, MINB, MINBS)                                                                  \
  , blind_rotate_kernel<LOGN, L, BG, SMALL, MINB, true>, blind_rotate_staged_kernel<LOGN, L, BG, SMALL, MINBS>,        \
    br_staged_smem<LOGN>, w16_kernel<LOGN, L, BG, SMALL>(), tm_kernel<LOGN, L, BG, SMALL>(),                           \
    tx_kernel<LOGN, L, BG, SMALL>(), txs_kernel<LOGN, L, BG, SMALL>(), tms_kernel<LOGN, L, BG, SMALL>(), br_tms_smem<LOGN>, \
    mg_kernel<LOGN, L, BG, SMALL>(), br_mg_smem<LOGN>, lat2_kernel<LOGN, L, BG, SMALL>(), br_lat2_smem<LOGN, L>,        \
    cl_kernel<LOGN, L, BG, SMALL>(), br_cl_smem<LOGN>
#else
#define VARIANT_EXP(LOGN, L, BG, SMALL, MINB, MINBS)
#endif
// The throughput kernel is instantiated in blind_rotate_throughput.cu (its own ptxas options); this translation unit only
// asks for the kernel's address (a build with -DTFHE_BR_SINGLE_TU instantiates it here instead: tools/build_exp.sh).
#ifdef TFHE_BR_SINGLE_TU
#define TFHE_BR_THROUGHPUT_PTR(LOGN, L, BG, SMALL, MINB) blind_rotate_kernel<LOGN, L, BG, SMALL, MINB>
#else
#define TFHE_BR_THROUGHPUT_PTR(LOGN, L, BG, SMALL, MINB) blind_rotate_throughput_kernel(LOGN, L, BG)
#endif
#define VARIANT(LOGN, L, BG, SMALL, MINB, MINBS)                                                     \
  { LOGN, L, BG, SMALL, TFHE_BR_THROUGHPUT_PTR(LOGN, L, BG, SMALL, MINB), cmux_kernel<LOGN, L, BG, SMALL, MINB>, \
    br_smem<LOGN>, lat_kernel<LOGN, L, BG, SMALL>(), br_lat_smem<LOGN>, latp_kernel<LOGN, L, BG, SMALL>(),     \
    br_latp_smem<LOGN, L> VARIANT_EXP(LOGN, L, BG, SMALL, MINB, MINBS) }
const Variant kVariants[] = {
    VARIANT(10, 3, 6, true, TFHE_BR_MINB_N1024, TFHE_BR_STAGED_MINB_N1024),  // 80 / 110 / 128-bit (params/params.go:83-180)
    VARIANT(10, 2, 10, false, 4, 4),  // Uint1               (params.go:194-223)
    VARIANT(9, 1, 18, false, 8, 8),   // Uint2               (params.go:236-265)
    VARIANT(10, 1, 23, false, 4, 4),  // Uint3               (params.go:277-306)
    VARIANT(11, 1, 22, false, 2, 2),  // Uint4 / Uint5       (params.go:318-391)
};

int find_variant(const tfhe_params& P) {
  int logN = 0;
  while ((1 << logN) < P.N) logN++;
  if ((1 << logN) != P.N) return -1;
  for (size_t v = 0; v < sizeof(kVariants) / sizeof(kVariants[0]); v++)
    if (kVariants[v].logN == logN && kVariants[v].L == P.L && kVariants[v].bgbit == P.bgbit) return (int)v;
  return -1;
}

size_t cmux_smem(int N) { return (size_t)8 * N + (size_t)32 * N + (size_t)8 * N; }

int set_device(tfhe_ctx* c) {
  CK(c, cudaSetDevice(c->device));
  return 0;
}

// --- engine steps on device buffers --------------------------------------------------------------
// Work-item size of the persistent throughput kernel.  A batch that fits the resident blocks runs whole gates (nothing to
// balance).  Larger batches are cut into items of ~n/14 steps, the count in [10, 20] chosen so that the LAST round of
// items over the resident blocks is as full as possible (4096 gates at n = 700 over 592 blocks: 13 items of 54 steps =
// 89.95 rounds).
void pick_chunks(const tfhe_ctx* c, int64_t count, BrArgs* a) {
  const int n = c->P.n;
  const int64_t resident = (int64_t)c->sm_count * c->br_blocks_per_sm;
  a->nchunks = 1; a->chunk_steps = n; a->main_chunks = 1;
  for (int k = 0; k <= BR_MAX_TAIL; k++) a->tail_start[k] = n;
  if (c->br_chunk_steps > 0) {  // forced: uniform items
    a->chunk_steps = std::min(n, c->br_chunk_steps);
    a->nchunks = a->main_chunks = (n + a->chunk_steps - 1) / a->chunk_steps;
    return;
  }
  if (count <= resident || n < 64) return;
  double best = 1e300;
  int kk_best = 1, steps_best = n;
  for (int k = 10; k <= 20; k++) {
    const int steps = (n + k - 1) / k;
    const int kk = (n + steps - 1) / steps;
    const double rounds = (double)count * kk / (double)resident;
    const double cost = std::ceil(rounds) / rounds * (1.0 + 0.0005 * kk);  // last-round fill, small per-item overhead
    if (cost < best) { best = cost; kk_best = kk; steps_best = steps; }
  }
  // Graded tail: the blocks of a launch finish spread over one item's duration (half of it idle on average, ~0.2 ms with
  // 54-step items), so the LAST item of every gate is cut into pieces of half, a quarter, ... of its length: the launch
  // then ends within a few steps' time for the price of a handful of extra hand-overs per gate.
  a->chunk_steps = steps_best;
  a->main_chunks = kk_best - 1;
  int lo = a->main_chunks * steps_best, nt = 0;
  a->tail_start[0] = lo;
  while (lo < n && nt < BR_MAX_TAIL) {
    const int left = n - lo;
    const int piece = (nt == BR_MAX_TAIL - 1 || left <= c->br_tail_min) ? left : (left + 1) / 2;
    lo += piece;
    a->tail_start[++nt] = lo;
  }
  for (int k = nt + 1; k <= BR_MAX_TAIL; k++) a->tail_start[k] = n;
  a->nchunks = a->main_chunks + nt;
}

// options of a blind rotation beyond the reference's: LUT table + index, many-LUT mod switch, multi-index extraction
struct BrOpts {
  const int* d_lut_index = nullptr;  // [count] indices into d_luts
  int ms_log2k = 0;                  // mod switch keeps multiples of 2^ms_log2k
  int extract_k = 0;                 // out_mode 2: samples at indices 0..extract_k-1
};

int launch_blind_rotate(tfhe_ctx* c, int64_t count, const uint32_t* d_ct, const uint32_t* d_luts, int64_t nluts,
                        uint32_t* d_out, int out_mode, cudaStream_t s, const BrOpts& opt = BrOpts()) {
  if (count == 0) return 0;
  const Variant& V = kVariants[c->variant];
  BrArgs a{};
  a.ct_in = d_ct; a.testvec = c->d_testvec; a.luts = d_luts; a.nluts = nluts; a.bsk = c->d_bsk; a.tw_tab = c->d_tw;
  a.out = d_out; a.n = c->P.n; a.offset = c->offset; a.out_mode = out_mode; a.tw0 = c->tw0;
  a.lut_index = opt.d_lut_index; a.ms_log2k = opt.ms_log2k; a.extract_k = opt.extract_k;
  a.count = count; a.nchunks = 1; a.chunk_steps = c->P.n; a.main_chunks = 1;
  for (int k = 0; k <= BR_MAX_TAIL; k++) a.tail_start[k] = c->P.n;
  const int T = c->P.N / 16;
#if TFHE_EXPERIMENTAL
  if (c->br_variant == 3 && V.br_w16 && c->d_bsk16) {
    BrW16Args w{};
    w.ct_in = d_ct; w.testvec = c->d_testvec; w.luts = d_luts; w.nluts = nluts; w.count = count; w.bsk = c->d_bsk16;
    w.tw1 = c->d_tw16; w.twl = c->d_tw16 + 8 * 16; w.out = d_out; w.n = c->P.n; w.offset = c->offset; w.out_mode = out_mode;
    w.tw0 = c->tw0_16;
    V.br_w16<<<(unsigned)((count + 3) / 4), 128, br_w16_smem_bytes(c->P.n), s>>>(w);
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
  }
  a.bsk_tex = c->bsk_tex;
  bool done = true;
  if (c->br_variant == 6 && V.br_txs) V.br_txs<<<(unsigned)count, T, V.br_staged_smem(c->P.n), s>>>(a);
  else if (c->br_variant == 5 && V.br_tx) V.br_tx<<<(unsigned)count, T, V.br_smem(c->P.n), s>>>(a);
  else if (V.br_cl && c->br_variant == 13) V.br_cl<<<(unsigned)(count * 2 * c->P.L), T, V.br_cl_smem(c->P.n), s>>>(a);
  else if (V.br_lat2 && c->br_variant == 11) V.br_lat2<<<(unsigned)count, 2 * c->P.L * T, V.br_lat2_smem(c->P.n), s>>>(a);
  else if (c->br_variant == 8 && V.br_mg)
    V.br_mg<<<(unsigned)((count + TFHE_BR_MG_G - 1) / TFHE_BR_MG_G), TFHE_BR_MG_G * T, V.br_mg_smem(c->P.n), s>>>(a, (long long)count);
  else if (c->br_variant == 7 && V.br_tms) V.br_tms<<<(unsigned)count, T, V.br_tms_smem(c->P.n), s>>>(a);
  else if (c->br_variant == 4 && V.br_tm) V.br_tm<<<(unsigned)count, T, V.br_smem(c->P.n), s>>>(a);
  else if (c->br_variant == 1) V.br_staged<<<(unsigned)count, T, V.br_staged_smem(c->P.n), s>>>(a);
  else done = false;
  if (done) {
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
  }
#endif
  // order-preserving latency mode (Uint / PBS sets): a batch of at most one ciphertext per SM
  if (V.br_latp && (c->br_variant == 12 || (c->br_variant == 0 && c->br_auto_lat && count <= (int64_t)c->sm_count))) {
    V.br_latp<<<(unsigned)count, 2 * T, V.br_latp_smem(c->P.n), s>>>(a);
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
  }
  // latency mode: a batch that cannot fill the SMs with the throughput kernel (<= 2 gates per SM) gets four warps per gate
  if (V.br_lat && (c->br_variant == 9 || (c->br_variant == 0 && c->br_auto_lat && count <= 2 * (int64_t)c->sm_count))) {
    V.br_lat<<<(unsigned)count, 2 * T, V.br_lat_smem(c->P.n), s>>>(a);
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
  }
  // throughput kernel: persistent blocks over work items, in sub-batches that bound the hand-over buffer
  const int64_t SUB = 32768;
  const int64_t resident = (int64_t)c->sm_count * c->br_blocks_per_sm;
  for (int64_t g0 = 0; g0 < count; g0 += SUB) {
    const int64_t cnt = std::min<int64_t>(SUB, count - g0);
    BrArgs b = a;
    b.count = cnt;
    b.ct_in = d_ct + (size_t)g0 * (c->P.n + 1);
    if (d_luts && nluts != 1 && !opt.d_lut_index) b.luts = d_luts + (size_t)g0 * 2 * c->P.N;
    if (opt.d_lut_index) b.lut_index = opt.d_lut_index + g0;
    b.out = d_out + (size_t)g0 * (out_mode == 0 ? 2 * c->P.N : (out_mode == 2 ? opt.extract_k : 1) * (c->P.N + 1));
    pick_chunks(c, cnt, &b);
    const size_t ctl_bytes = 16 + (size_t)std::min<int64_t>(SUB, std::max<int64_t>(cnt, 1)) * sizeof(int);
    if (ctl_bytes > c->br_ctl.cap) {  // (re)allocated control words start at zero; the kernel leaves them at zero
      CK(c, c->br_ctl.reserve(16 + (size_t)SUB * sizeof(int)));
      CK(c, cudaMemsetAsync(c->br_ctl.p, 0, c->br_ctl.cap, s));
    }
    b.ctl = c->br_ctl.as<unsigned int>();
    b.progress = reinterpret_cast<int*>(c->br_ctl.as<unsigned char>() + 16);
    if (b.nchunks > 1) {
      if (out_mode == 0) b.scratch = b.out;  // the TRLWE output rows double as the hand-over buffer
      else {
        CK(c, c->br_scratch.reserve((size_t)cnt * 2 * c->P.N * 4));
        b.scratch = c->br_scratch.as<uint32_t>();
      }
    }
    const int64_t items = cnt * b.nchunks;
    const unsigned grid = (unsigned)std::min<int64_t>(items, resident);
#if TFHE_EXPERIMENTAL
    if (c->br_variant == 2) V.br_tex<<<grid, T, V.br_smem(c->P.n), s>>>(b);
    else
#endif
    V.br<<<grid, T, V.br_smem(c->P.n), s>>>(b);
    c->launches++;
    CK(c, cudaGetLastError());
  }
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// 2-D u8 tensor [rows][K], K contiguous; box = 128 K-bytes x box_rows, 128-byte swizzle, out-of-range rows read as zero
int make_ks_map(tfhe_ctx* c, CUtensorMap* m, const void* base, long long K, long long rows, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(c, TFHE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K};
  const cuuint32_t box[2] = {(cuuint32_t)KSM_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, TFHE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// [rows][stride] u32 matrix, box = box_rows x box_words, no swizzle, out-of-range columns read as zeros
int make_u32_map(tfhe_ctx* c, CUtensorMap* m, const void* base, long long stride_words, long long rows, int box_words, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(c, TFHE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {(cuuint64_t)stride_words, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)stride_words * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_words, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(c, TFHE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// Tensor-core path: chunks of <= 8192 ciphertexts (selection matrix <= K * 8192 bytes).
int launch_key_switch_mma(tfhe_ctx* c, int64_t count, const uint32_t* d_lwe1, uint32_t* d_out, cudaStream_t s,
                          const GateDesc* out_gates, long long instances) {
  const tfhe_params& P = c->P;
  const long long K = c->ks_K;
  const int64_t CH = 8192;
  CK(c, c->ks_sel.reserve((size_t)std::min<int64_t>(count, CH) * K));
  const int cols = 4 * c->ksk_stride;
  for (int64_t g0 = 0; g0 < count; g0 += CH) {
    const int cnt = (int)std::min<int64_t>(CH, count - g0);
    const uint32_t* src = d_lwe1 + (size_t)g0 * (P.N + 1);
    ks_onehot_kernel<<<(unsigned)cnt, 256, (size_t)K, s>>>(src, c->ks_sel.as<uint8_t>(), d_out, P.N, P.n, P.basebit, P.iks_t,
                                                             (int)K, out_gates, instances, (long long)g0);
    c->launches++;
    CK(c, cudaGetLastError());
    CUtensorMap mapA;
    int rc = make_ks_map(c, &mapA, c->ks_sel.p, K, cnt, KSM_BM);
    if (rc) return rc;
    KsMmaArgs a{};
    a.out = d_out; a.out_gates = out_gates; a.instances = instances; a.g_base = (long long)g0; a.count = cnt; a.n = P.n;
    a.kblocks = (int)(K / KSM_BK);
    a.tiles_m = (cnt + KSM_BM - 1) / KSM_BM;
    a.tiles_n = (cols + KSM_BN - 1) / KSM_BN;
    // split K so that the work items fill whole waves of SMs; ~8 k-blocks of fixed cost per item (setup + epilogue)
    const int tiles = a.tiles_m * a.tiles_n;
    int best = 1;
    double best_cost = 1e300;
    for (int ks = 1; ks <= 16 && ks <= a.kblocks; ks++) {
      const double waves = std::ceil((double)tiles * ks / c->sm_count);
      const double cost = waves * ((double)a.kblocks / ks + 8.0);
      if (cost < best_cost) { best_cost = cost; best = ks; }
    }
    a.ksplit = best;
    ks_mma_kernel<<<(unsigned)(tiles * a.ksplit), KSM_THREADS, KSM_SMEM, s>>>(mapA, c->ks_mapB, a);
    c->launches++;
    CK(c, cudaGetLastError());
  }
  return 0;
}

int launch_key_switch_tile(tfhe_ctx* c, int64_t count, const uint32_t* d_lwe1, uint32_t* d_out, cudaStream_t s,
                           const GateDesc* out_gates, long long instances) {
  const auto& P = c->P;
  const int base = 1 << P.basebit, K = P.N * P.iks_t;
  const int col_tiles = (P.n + 1 + KST_COLS - 1) / KST_COLS;
  const int stages = K / (KST_STAGE_ROWS / base);   // a stage = 128 key rows = 128 / base pairs (N * t is a multiple of 8)
  const int64_t CH = 65536;                          // ciphertexts per pass: bounds the digit matrix (K bytes per ciphertext)
  CK(c, c->ks_sel.reserve((size_t)K * ((std::min<int64_t>(count, CH) + KST_CT - 1) / KST_CT * KST_CT)));
  auto kern = base == 64 ? ks_tile_kernel<64> : base == 32 ? ks_tile_kernel<32> : ks_tile_kernel<16>;
  for (int64_t g0 = 0; g0 < count; g0 += CH) {
    const int64_t cnt = std::min<int64_t>(CH, count - g0);
    const int64_t ct_tiles = (cnt + KST_CT - 1) / KST_CT, cpad = ct_tiles * KST_CT;
    const uint32_t* src = d_lwe1 + (size_t)g0 * (P.N + 1);
    ks_digits_kernel<<<dim3((unsigned)ct_tiles, (unsigned)(P.N / 8)), 256, 0, s>>>(src, c->ks_sel.as<uint8_t>(), d_out, cnt, cpad, P.N, P.n,
                                                                                 P.basebit, P.iks_t, out_gates, instances, (long long)g0);
    c->launches++;
    // split the stages of a tile over `ksplit` blocks so that the blocks fill whole rounds of the resident slots (2 per
    // SM); each block pays ~12 stages' worth of prologue + epilogue
    const int64_t tiles = ct_tiles * col_tiles, slots = 2 * (int64_t)c->sm_count;
    int best = 1;
    double best_cost = 1e300;
    for (int ks = 1; ks <= 64 && ks * 32 <= stages; ks++) {
      const double rounds = std::ceil((double)tiles * ks / slots);
      const double cost = rounds * ((double)stages / ks + 12.0);
      if (cost < best_cost) { best_cost = cost; best = ks; }
    }
    kern<<<(unsigned)(tiles * best), KST_THREADS, kst_smem_bytes(base), s>>>(c->ks_tile_map, c->ks_sel.as<uint8_t>(), d_out, stages, cnt, cpad, P.n,
                                                                         col_tiles, best, (int)ct_tiles, out_gates, instances, (long long)g0);
    c->launches++;
    CK(c, cudaGetLastError());
  }
  return 0;
}

int launch_key_switch(tfhe_ctx* c, int64_t count, const uint32_t* d_lwe1, uint32_t* d_out, cudaStream_t s,
                      const GateDesc* out_gates = nullptr, long long instances = 1) {
  if (count == 0) return 0;
  if (c->ks_variant == 2 && !c->ks_K) return fail(c, TFHE_ERR_STATE, "tensor-core key switch is not available for this parameter set");
  // measured on B200 at 128-bit (host call incl. copies): 8 ciphertexts 0.07 vs 1.07 ms, 256: 0.20 vs 0.89 ms, 4096: 2.35 vs
  // 6.43 ms — the contraction wins at every batch size (a lone gather block streams its 19 MB at one SM's bandwidth)
  if (c->ks_K && (c->ks_variant == 2 || c->ks_variant == 0))
    return launch_key_switch_mma(c, count, d_lwe1, d_out, s, out_gates, instances);
  // large-base sets (Uint2-5), a tile's worth of ciphertexts or more: shared-memory tiles (key_switch_tile.cuh).  One tile of
  // 256 ciphertexts costs what ~140 gathered ciphertexts cost (Uint5), so the switch-over sits below a full tile.
  if (c->ks_tile_ok && (c->ks_variant == 3 || (c->ks_variant == 0 && count >= 160)))
    return launch_key_switch_tile(c, count, d_lwe1, d_out, s, out_gates, instances);
  if (c->ks_variant == 3) return fail(c, TFHE_ERR_STATE, "tiled key switch is for the basebit 4..6 parameter sets");
  const size_t sm = (size_t)c->P.N * c->P.iks_t * sizeof(uint32_t);
  if (sm > 128 * 1024) return fail(c, TFHE_ERR_ARG, "N * iks_t too large for the key-switch kernel");
  // one thread per 16-byte column of a key row, so that the row loop runs once (n = 1071: 268 columns -> 288 threads,
  // not 256 + a second pass with 12 live threads)
  const int ks_threads = std::min(512, std::max(128, (c->ksk_stride / 4 + 31) / 32 * 32));
  // fewer ciphertexts than ~2 blocks per SM: split each ciphertext's rows over several blocks
  const int splits = (int)std::max<int64_t>(1, std::min<int64_t>(64, (2 * (int64_t)c->sm_count) / count));
  if (splits > 1) {
    zero_out_rows_kernel<<<(unsigned)count, 128, 0, s>>>(d_out, c->P.n, out_gates, instances);
    c->launches++;
  }
  key_switch_kernel<<<dim3((unsigned)count, (unsigned)splits), ks_threads, sm, s>>>(d_lwe1, c->d_ksk, d_out, c->P.N, c->P.n, c->P.basebit,
                                                                               c->P.iks_t, c->ksk_stride, out_gates, instances, splits);
  c->launches++;
  CK(c, cudaGetLastError());
  return 0;
}

bool is_group(const tfhe_ctx* c) { return !c->kids.empty(); }

// a setting applied to a group goes to every device
#define GROUP_FORWARD(c, call)                                                   \
  if ((c) && is_group(c)) {                                                      \
    for (tfhe_ctx* k__ : (c)->kids) {                                            \
      tfhe_ctx* kid = k__;                                                       \
      const int rc__ = (call);                                                   \
      if (rc__) { (c)->err = kid->err; return rc__; }                            \
    }                                                                            \
    return TFHE_OK;                                                              \
  }


// --- multi-device groups ------------------------------------------------------------------------------------------
// Contiguous shards of [0, count) over the kids of a group, cut where the running COST (bootstraps) is closest to an even
// split; cost == nullptr: equal counts.
std::vector<int64_t> shard_bounds(int ndev, int64_t count, const uint8_t* ops, int64_t nops) {
  std::vector<int64_t> b(ndev + 1, 0);
  if (!ops || nops == 1) {
    for (int d = 0; d <= ndev; d++) b[d] = count * d / ndev;
    return b;
  }
  auto cost = [](uint8_t op) -> int64_t { return op == TFHE_OP_MUX ? 3 : (op >= TFHE_OP_NOT ? 0 : 1); };
  int64_t total = 0;
  for (int64_t g = 0; g < count; g++) total += cost(ops[g]);
  int64_t run = 0, g = 0;
  for (int d = 1; d < ndev; d++) {
    const int64_t want = total * d / ndev;
    while (g < count && run + cost(ops[g]) <= want) run += cost(ops[g++]);
    b[d] = g;
  }
  b[ndev] = count;
  return b;
}

// runs fn(kid index) for every kid, kid 0 on the calling thread; returns the first failure and copies its message
template <class Fn>
int for_each_kid(tfhe_ctx* c, Fn fn) {
  const int nd = (int)c->kids.size();
  std::vector<int> rc(nd, 0);
  std::vector<std::thread> th;
  for (int d = 1; d < nd; d++) th.emplace_back([&, d] { rc[d] = fn(d); });
  rc[0] = fn(0);
  for (auto& t : th) t.join();
  for (int d = 0; d < nd; d++)
    if (rc[d]) { c->err = "device " + std::to_string(c->kids[d]->device) + ": " + c->kids[d]->err; return rc[d]; }
  return TFHE_OK;
}


int check_ready(tfhe_ctx* c, bool need_ksk) {
  if (!c) return TFHE_ERR_ARG;
  if (!c->key_loaded) return fail(c, TFHE_ERR_STATE, "cloud key not loaded");
  if (need_ksk && !c->has_ksk) return fail(c, TFHE_ERR_STATE, "key-switching key not loaded");
  return 0;
}

int bootstrap_device(tfhe_ctx* c, int64_t count, const uint32_t* d_ct, const uint32_t* d_luts, int64_t nluts,
                     uint32_t* d_out, cudaStream_t s, const GateDesc* out_gates = nullptr, long long instances = 1,
                     const BrOpts& opt = BrOpts()) {
  const int64_t per = opt.extract_k > 0 ? opt.extract_k : 1;  // extracted samples (= key switches) per ciphertext
  CK(c, c->lwe1.reserve((size_t)count * per * (c->P.N + 1) * 4));
  tfhe_ctx::StageEv ev{};
  if (c->timing && count > 0) {
    if (!c->ev_free.empty()) { ev = c->ev_free.back(); c->ev_free.pop_back(); }
    else { CK(c, cudaEventCreate(&ev.e0)); CK(c, cudaEventCreate(&ev.e1)); CK(c, cudaEventCreate(&ev.e2)); }
    CK(c, cudaEventRecord(ev.e0, s));
  }
  int rc = launch_blind_rotate(c, count, d_ct, d_luts, nluts, c->lwe1.as<uint32_t>(), opt.extract_k > 0 ? 2 : 1, s, opt);
  if (rc) return rc;
  if (c->timing && count > 0) CK(c, cudaEventRecord(ev.e1, s));
  rc = launch_key_switch(c, count * per, c->lwe1.as<uint32_t>(), d_out, s, out_gates, instances);
  if (rc) return rc;
  if (c->timing && count > 0) { CK(c, cudaEventRecord(ev.e2, s)); c->ev_live.push_back(ev); }
  return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* tfhe_version(void) { return "tfhe_b200 0.1 (sm_100a)"; }

const char* tfhe_last_error(const tfhe_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  std::lock_guard<std::mutex> l(g_create_mu);
  return g_create_error.c_str();
}

int tfhe_ctx_create(const tfhe_params* params, int device, tfhe_ctx** out) {
  if (!params || !out) return fail(nullptr, TFHE_ERR_ARG, "null argument");
  *out = nullptr;
  const tfhe_params& P = *params;
  if (P.n <= 0 || P.n > 4096 || P.L <= 0 || P.bgbit <= 0 || P.L * P.bgbit > 32 || P.basebit <= 0 || P.iks_t <= 0 ||
      P.basebit * P.iks_t > 31)
    return fail(nullptr, TFHE_ERR_ARG, "invalid parameters");
  const int v = find_variant(P);
  if (v < 0)
    return fail(nullptr, TFHE_ERR_ARG, "unsupported shape N=%d L=%d bgbit=%d (supported: the reference's 80/110/128-bit and Uint1-5 sets)",
                P.N, P.L, P.bgbit);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, TFHE_ERR_CUDA, "no CUDA device (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, TFHE_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, TFHE_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, TFHE_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                prop.minor);
  tfhe_ctx* c = new (std::nothrow) tfhe_ctx();
  if (!c) return fail(nullptr, TFHE_ERR_NOMEM, "out of host memory");
  c->P = P; c->device = device; c->variant = v; c->logN = kVariants[v].logN; c->sm_count = prop.multiProcessorCount;
  auto bail = [&](const char* what, cudaError_t err) {
    int rc = fail(nullptr, TFHE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
    tfhe_ctx_destroy(c);
    return rc;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (Slot& sl : c->slot)
    for (cudaEvent_t* ev : {&sl.in_done, &sl.comp_done, &sl.out_done, &sl.idx_done})
      if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  if (const char* pc = getenv("TFHE_B200_PIPE_CHUNK")) c->pipe_chunk = std::max<int64_t>(256, atoll(pc));
  std::vector<Tw4> tab;
  switch (c->logN) {
    case 9: build_twiddles<8>(c->tw0, tab); break;
    case 10: build_twiddles<9>(c->tw0, tab); break;
    case 11: build_twiddles<10>(c->tw0, tab); break;
    default: tfhe_ctx_destroy(c); return fail(nullptr, TFHE_ERR_ARG, "unsupported N");
  }
#if TFHE_BR_TW_CONST
  if (tab.size() >= 8 && (e = cudaMemcpyToSymbol(c_tw_pass1, tab.data(), 8 * sizeof(Tw4), (size_t)(c->logN - 9) * 8 * sizeof(Tw4))) != cudaSuccess)
    return bail("cudaMemcpyToSymbol(twiddles)", e);
#endif
  if ((e = cudaMalloc(&c->d_tw, tab.size() * sizeof(Tw4))) != cudaSuccess) return bail("cudaMalloc(twiddles)", e);
  if ((e = cudaMemcpy(c->d_tw, tab.data(), tab.size() * sizeof(Tw4), cudaMemcpyHostToDevice)) != cudaSuccess)
    return bail("cudaMemcpy(twiddles)", e);
  const Variant& V = kVariants[v];
  if (V.br_latp && (e = cudaFuncSetAttribute(V.br_latp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_latp_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_latp)", e);
  if (V.br_lat && (e = cudaFuncSetAttribute(V.br_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_lat_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_lat)", e);
  if ((e = cudaFuncSetAttribute(V.br, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_smem(4096))) != cudaSuccess)  // limit, not allocation: n <= 4096
    return bail("cudaFuncSetAttribute(blind_rotate)", e);
  {  // resident blocks of the persistent throughput kernel (registers: 4 at N = 1024)
    int nb = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, V.br, P.N / 16, V.br_smem(P.n))) != cudaSuccess)
      return bail("cudaOccupancyMaxActiveBlocksPerMultiprocessor(blind_rotate)", e);
    c->br_blocks_per_sm = nb > 0 ? nb : 1;
  }
  if (const char* cs = getenv("TFHE_B200_BR_CHUNK_STEPS")) c->br_chunk_steps = atoi(cs);
  if (const char* cs = getenv("TFHE_B200_BR_TAIL_MIN")) c->br_tail_min = std::max(1, atoi(cs));
  if (const char* sel = getenv("TFHE_B200_BR"))
    c->br_variant = !strcmp(sel, "lat") ? 9 : !strcmp(sel, "latp") ? 12 : !strcmp(sel, "throughput") ? 10 : !strcmp(sel, "ldg") ? 0 : c->br_variant;
#if TFHE_EXPERIMENTAL
  if (V.br_cl && (e = cudaFuncSetAttribute(V.br_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_cl_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_cl)", e);
  if (V.br_lat2 && (e = cudaFuncSetAttribute(V.br_lat2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_lat2_smem(2048))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_lat2)", e);
  if (V.br_mg && (e = cudaFuncSetAttribute(V.br_mg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_mg_smem(2048))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_mg)", e);
  if (V.br_tms && (e = cudaFuncSetAttribute(V.br_tms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_tms_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_tms)", e);
  if (V.br_tm && (e = cudaFuncSetAttribute(V.br_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_tm)", e);
  if (V.br_tx && (e = cudaFuncSetAttribute(V.br_tx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_tx)", e);
  if (V.br_txs && (e = cudaFuncSetAttribute(V.br_txs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_staged_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_txs)", e);
  if (V.br_w16) {
    const int M = 512;
    auto blk = [&](int m, int i, double2* o) {
      o[0] = twiddle(M, m, i); o[1] = twiddle(M, 2 * m, 2 * i); o[2] = twiddle(M, 4 * m, 4 * i); o[3] = twiddle(M, 4 * m, 4 * i + 2);
      for (int j = 0; j < 4; j++) o[4 + j] = twiddle(M, 8 * m, 8 * i + 2 * j);
    };
    blk(1, 0, c->tw0_16.s);
    std::vector<double2> t16(8 * 16 + 4 * 32);
    for (int b = 0; b < 16; b++) {
      double2 o[8];
      blk(16, b, o);
      for (int j = 0; j < 8; j++) t16[j * 16 + b] = o[j];
    }
    for (int lane = 0; lane < 32; lane++)
      for (int j = 0; j < 4; j++) t16[8 * 16 + j * 32 + lane] = twiddle(M, 256, 8 * lane + 2 * j);
    if ((e = cudaMalloc(&c->d_tw16, t16.size() * sizeof(double2))) != cudaSuccess) return bail("cudaMalloc(twiddles16)", e);
    if ((e = cudaMemcpy(c->d_tw16, t16.data(), t16.size() * sizeof(double2), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail("cudaMemcpy(twiddles16)", e);
    if ((e = cudaFuncSetAttribute(V.br_w16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_w16_smem_bytes(4096))) != cudaSuccess)
      return bail("cudaFuncSetAttribute(blind_rotate_w16)", e);
  }
  if ((e = cudaFuncSetAttribute(V.br_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_staged_smem(4096))) !=
      cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_staged)", e);
  if ((e = cudaFuncSetAttribute(V.br_tex, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V.br_smem(4096))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(blind_rotate_tex)", e);
  if (const char* sel = getenv("TFHE_B200_BR"))
    c->br_variant = !strcmp(sel, "tma") ? 1 : !strcmp(sel, "tex") ? 2 : !strcmp(sel, "w16") ? 3 : !strcmp(sel, "tmem") ? 4 : !strcmp(sel, "tmex") ? 5 : !strcmp(sel, "tmex+tma") ? 6 : !strcmp(sel, "tms") ? 7 : !strcmp(sel, "mg") ? 8 : !strcmp(sel, "lat2") ? 11 : !strcmp(sel, "cl") ? 13 : c->br_variant;
#endif
  if (c->br_variant == 10) { c->br_variant = 0; c->br_auto_lat = false; }
  {  // key generation: exchange buffers + the row's mask words (56 KiB at N = 2048: above the 48 KiB default)
    const int kg_smem = (int)((size_t)br_nbuf(c->logN) * TFHE_BR_EXW * (P.N / 2) * 16 + (size_t)P.N * 4);
    e = c->logN == 9 ? cudaFuncSetAttribute(keygen_bsk_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, kg_smem)
      : c->logN == 10 ? cudaFuncSetAttribute(keygen_bsk_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, kg_smem)
                      : cudaFuncSetAttribute(keygen_bsk_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, kg_smem);
    if (e != cudaSuccess) return bail("cudaFuncSetAttribute(keygen_bsk)", e);
  }
  if ((e = cudaFuncSetAttribute(V.cmux, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cmux_smem(P.N))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(cmux)", e);
  if ((e = cudaFuncSetAttribute(ks_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KSM_SMEM)) != cudaSuccess)
    return bail("cudaFuncSetAttribute(ks_mma)", e);
  if ((e = cudaFuncSetAttribute(ks_onehot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess)
    return bail("cudaFuncSetAttribute(ks_onehot)", e);
  if (P.basebit >= 4 && P.basebit <= 6 &&
      (e = cudaFuncSetAttribute(P.basebit == 6 ? ks_tile_kernel<64> : P.basebit == 5 ? ks_tile_kernel<32> : ks_tile_kernel<16>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kst_smem_bytes(1 << P.basebit))) != cudaSuccess)
    return bail("cudaFuncSetAttribute(ks_tile)", e);
  if ((e = cudaFuncSetAttribute(key_switch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                128 * 1024)) != cudaSuccess)  // a limit shared by every context of the process
    return bail("cudaFuncSetAttribute(key_switch)", e);
  *out = c;
  return TFHE_OK;
}

// One context over several GPUs of this process: the batch entry points with host buffers (tfhe_gate_batch,
// tfhe_bootstrap_batch, tfhe_blind_rotate_batch, tfhe_circuit_run) shard their batch over the devices — what
// trgsw.BatchBlindRotate's goroutine fan-out (trgsw/trgsw.go:234-252) becomes on a multi-GPU node.
int tfhe_ctx_create_multi(const tfhe_params* params, int ndev, const int* devices, tfhe_ctx** out) {
  if (!params || !out) return fail(nullptr, TFHE_ERR_ARG, "null argument");
  *out = nullptr;
  int visible = 0;
  cudaError_t e = cudaGetDeviceCount(&visible);
  if (e != cudaSuccess || visible == 0)
    return fail(nullptr, TFHE_ERR_CUDA, "no CUDA device (%s); this engine has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (ndev <= 0) { ndev = visible; devices = nullptr; }
  if (ndev > visible && !devices) return fail(nullptr, TFHE_ERR_ARG, "%d devices requested, %d visible", ndev, visible);
  tfhe_ctx* g = new (std::nothrow) tfhe_ctx();
  if (!g) return fail(nullptr, TFHE_ERR_NOMEM, "out of host memory");
  g->P = *params;
  for (int d = 0; d < ndev; d++) {
    const int dev = devices ? devices[d] : d;
    for (tfhe_ctx* k : g->kids)
      if (k->device == dev) { tfhe_ctx_destroy(g); return fail(nullptr, TFHE_ERR_ARG, "device %d listed twice", dev); }
    tfhe_ctx* k = nullptr;
    const int rc = tfhe_ctx_create(params, dev, &k);
    if (rc) {
      if (g->kids.empty()) delete g; else tfhe_ctx_destroy(g);
      return rc;  // message already recorded by tfhe_ctx_create
    }
    g->kids.push_back(k);
  }
  g->device = g->kids[0]->device; g->variant = g->kids[0]->variant; g->logN = g->kids[0]->logN; g->sm_count = g->kids[0]->sm_count;
  *out = g;
  return TFHE_OK;
}

int tfhe_ctx_device_count(const tfhe_ctx* c) { return c ? (is_group(c) ? (int)c->kids.size() : 1) : 0; }

void tfhe_ctx_destroy(tfhe_ctx* c) {
  if (!c) return;
  if (is_group(c)) {
    for (tfhe_ctx* k : c->kids) tfhe_ctx_destroy(k);
    delete c;
    return;
  }
  cudaSetDevice(c->device);
  for (cudaStream_t* st : {&c->stream, &c->s_in, &c->s_out})
    if (*st) { cudaStreamSynchronize(*st); cudaStreamDestroy(*st); }
  for (Slot& sl : c->slot) {
    for (DevBuf* b : {&sl.a, &sl.b, &sl.c, &sl.luts, &sl.out}) b->release();
    for (cudaEvent_t ev : {sl.in_done, sl.comp_done, sl.out_done, sl.idx_done})
      if (ev) cudaEventDestroy(ev);
    if (sl.h_idx) cudaFreeHost(sl.h_idx);
  }
  for (DevBuf* b : {&c->wires, &c->gate_descs, &c->prep, &c->lwe1, &c->tmp, &c->prep2, &c->idx_a, &c->idx_b, &c->ops_dev, &c->h2d_a, &c->h2d_b,
                    &c->h2d_c, &c->h2d_luts, &c->d2h_out, &c->br_ctl, &c->br_scratch})
    b->release();
  for (auto& q : c->circuit_graphs)
    if (q.exec) cudaGraphExecDestroy(q.exec);
  c->circuit_graphs.clear();
  for (auto* v : {&c->ev_live, &c->ev_free})
    for (auto& ev : *v) { cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1); cudaEventDestroy(ev.e2); }
  if (c->bsk_tex) cudaDestroyTextureObject(c->bsk_tex);
  if (c->d_bsk) cudaFree(c->d_bsk);
#if TFHE_EXPERIMENTAL
  if (c->d_bsk16) cudaFree(c->d_bsk16);
  if (c->d_tw16) cudaFree(c->d_tw16);
#endif
  if (c->d_ksk) cudaFree(c->d_ksk);
  if (c->d_ksk_bytes) cudaFree(c->d_ksk_bytes);
  if (c->d_reenc) cudaFree(c->d_reenc);
  c->ks_sel.release();
  if (c->d_testvec) cudaFree(c->d_testvec);
  if (c->d_tw) cudaFree(c->d_tw);
  delete c;
}

int tfhe_ctx_load_cloudkey_device(tfhe_ctx* c, uint32_t offset, const double* d_bsk_fft, const uint32_t* d_ksk,
                                  const uint32_t* d_testvec, void* stream) {
  if (!c || !d_bsk_fft || !d_testvec) return fail(c, TFHE_ERR_ARG, "null argument");
  if (is_group(c)) return fail(c, TFHE_ERR_ARG, "device-buffer entry points need a single-device context");
  int rc = set_device(c);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const tfhe_params& P = c->P;
  const size_t polys = (size_t)P.n * 2 * P.L * 2;
  const int M = P.N / 2;
  if (!c->d_bsk) CK(c, cudaMalloc(&c->d_bsk, polys * M * sizeof(double2) + (size_t)8 * 2 * M * sizeof(double2)));  // + slack for key-row prefetches past the last step
  if (!c->d_testvec) CK(c, cudaMalloc(&c->d_testvec, (size_t)2 * P.N * 4));
  bsk_repack_kernel<<<(unsigned)polys, 128, 0, s>>>(d_bsk_fft, c->d_bsk, P.N, P.L, P.bgbit);
  c->launches++;
  CK(c, cudaGetLastError());
#if TFHE_EXPERIMENTAL
  if (kVariants[c->variant].br_w16) {
    if (!c->d_bsk16) CK(c, cudaMalloc(&c->d_bsk16, polys * M * sizeof(double2)));
    bsk_repack_w16_kernel<<<(unsigned)polys, 128, 0, s>>>(d_bsk_fft, c->d_bsk16);
    c->launches++;
    CK(c, cudaGetLastError());
  }
#endif
  CK(c, cudaMemcpyAsync(c->d_testvec, d_testvec, (size_t)2 * P.N * 4, cudaMemcpyDeviceToDevice, s));
  c->has_ksk = false;  // a key loaded without a key-switching key must not be paired with the previous one
  c->ks_K = 0;
  if (d_ksk) {
    const size_t rows = (size_t)P.N * P.iks_t * (1u << P.basebit);
    c->ksk_stride = (P.n + 1 + 3) / 4 * 4;
    if (!c->d_ksk) CK(c, cudaMalloc(&c->d_ksk, rows * c->ksk_stride * 4));
    ksk_repack_kernel<<<(unsigned)rows, 128, 0, s>>>(d_ksk, c->d_ksk, P.n + 1, c->ksk_stride, 1 << P.basebit);
    c->launches++;
    CK(c, cudaGetLastError());
    c->has_ksk = true;
    c->ks_tile_ok = false;
    if (P.basebit >= 4 && P.basebit <= 6 && (P.N * P.iks_t) % (KST_STAGE_ROWS >> P.basebit) == 0 && encode_tiled_fn()) {
      int rcm = make_u32_map(c, &c->ks_tile_map, c->d_ksk, c->ksk_stride, (long long)rows, KST_COLS, KST_STAGE_ROWS);
      if (rcm) return rcm;
      c->ks_tile_ok = true;
    }
    // byte planes for the tensor-core key switch: only where the dense product is cheap (base - 1 = 3 rows per digit)
    const long long K = (long long)P.N * P.iks_t * ((1 << P.basebit) - 1);
    if (P.basebit == 2 && K % KSM_BK == 0 && K <= 160 * 1024 && encode_tiled_fn()) {
      const int cols = 4 * c->ksk_stride;
      if (!c->d_ksk_bytes) CK(c, cudaMalloc(&c->d_ksk_bytes, (size_t)cols * K));
      dim3 grid((unsigned)((K + 63) / 64), (unsigned)((c->ksk_stride + 63) / 64));
      ksk_bytes_repack_kernel<<<grid, 256, 0, s>>>(c->d_ksk, c->d_ksk_bytes, c->ksk_stride, P.basebit, K);
      c->launches++;
      CK(c, cudaGetLastError());
      int rcm = make_ks_map(c, &c->ks_mapB, c->d_ksk_bytes, K, cols, KSM_BN);
      if (rcm) return rcm;
      c->ks_K = K;
    }
  }
  CK(c, cudaStreamSynchronize(s));
#if TFHE_EXPERIMENTAL
  if (!c->bsk_tex) {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = c->d_bsk;
    rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
    rd.res.linear.sizeInBytes = polys * M * sizeof(double2);
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    if (cudaCreateTextureObject(&c->bsk_tex, &rd, &td, nullptr) != cudaSuccess) { c->bsk_tex = 0; cudaGetLastError(); }
  }
  if (c->br_variant == 2 && !c->bsk_tex) return fail(c, TFHE_ERR_CUDA, "texture object over the bootstrapping key failed");
#endif
  c->offset = offset;
  c->key_loaded = true;
  return TFHE_OK;
}

}  // extern "C"

namespace {
// Replicates a cloud key that sits in reference layout in device 0's staging buffers to every other device of a group:
// one peer copy per buffer (NVLink / NVSwitch when peer access exists, the driver's fallback otherwise), then the
// ordinary repack on each device.  This is the "one broadcast of the cloud key at init" of the multi-GPU design; nothing
// crosses devices afterwards.
int group_replicate(tfhe_ctx* g, uint32_t offset, const void* d_bsk, size_t bsk_bytes, const void* d_ksk, size_t ksk_bytes,
                    const void* d_tv, size_t tv_bytes) {
  const int dev0 = g->kids[0]->device;
  int rc = for_each_kid(g, [&](int d) -> int {
    if (d == 0) return 0;
    tfhe_ctx* k = g->kids[d];
    int r = set_device(k);
    if (r) return r;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, k->device, dev0) == cudaSuccess && can) {
      cudaError_t pe = cudaDeviceEnablePeerAccess(dev0, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else cudaGetLastError();
    }
    DevBuf b2, k2, t2;
    cudaError_t e = b2.reserve(bsk_bytes);
    if (e == cudaSuccess && d_ksk) e = k2.reserve(ksk_bytes);
    if (e == cudaSuccess) e = t2.reserve(tv_bytes);
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(b2.p, k->device, d_bsk, dev0, bsk_bytes, k->stream);
    if (e == cudaSuccess && d_ksk) e = cudaMemcpyPeerAsync(k2.p, k->device, d_ksk, dev0, ksk_bytes, k->stream);
    if (e == cudaSuccess) e = cudaMemcpyPeerAsync(t2.p, k->device, d_tv, dev0, tv_bytes, k->stream);
    if (e == cudaSuccess)
      r = tfhe_ctx_load_cloudkey_device(k, offset, b2.as<double>(), d_ksk ? k2.as<uint32_t>() : nullptr, t2.as<uint32_t>(), k->stream);
    b2.release(); k2.release(); t2.release();
    if (e != cudaSuccess) return fail(k, TFHE_ERR_CUDA, "peer copy of the cloud key: %s", cudaGetErrorString(e));
    return r;
  });
  if (rc == 0) { g->key_loaded = true; g->has_ksk = d_ksk != nullptr; g->offset = offset; }
  return rc;
}
}  // namespace

extern "C" {

int tfhe_ctx_load_cloudkey(tfhe_ctx* g, uint32_t offset, const double* bsk_fft, const uint32_t* ksk,
                           const uint32_t* testvec) {
  if (!g || !bsk_fft || !testvec) return fail(g, TFHE_ERR_ARG, "null argument");
  tfhe_ctx* c = is_group(g) ? g->kids[0] : g;
  int rc = set_device(c);
  if (rc) return rc;
  const tfhe_params& P = c->P;
  const size_t bsk_bytes = (size_t)P.n * 2 * P.L * 2 * P.N * sizeof(double);
  const size_t ksk_bytes = ksk ? (size_t)P.N * P.iks_t * (1u << P.basebit) * (P.n + 1) * 4 : 0;
  const size_t tv_bytes = (size_t)2 * P.N * 4;
  // stage the reference-layout key in device memory, repack, then drop the staging copy
  DevBuf sb, sk, st;
  CK(c, sb.reserve(bsk_bytes));
  cudaError_t e = cudaMemcpy(sb.p, bsk_fft, bsk_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && ksk) { e = sk.reserve(ksk_bytes); if (e == cudaSuccess) e = cudaMemcpy(sk.p, ksk, ksk_bytes, cudaMemcpyHostToDevice); }
  if (e == cudaSuccess) { e = st.reserve(tv_bytes); if (e == cudaSuccess) e = cudaMemcpy(st.p, testvec, tv_bytes, cudaMemcpyHostToDevice); }
  if (e == cudaSuccess)
    rc = tfhe_ctx_load_cloudkey_device(c, offset, sb.as<double>(), ksk ? sk.as<uint32_t>() : nullptr, st.as<uint32_t>(),
                                       c->stream);
  if (e == cudaSuccess && rc == 0 && is_group(g))
    rc = group_replicate(g, offset, sb.p, bsk_bytes, ksk ? sk.p : nullptr, ksk_bytes, st.p, tv_bytes);
  set_device(c);
  sb.release(); sk.release(); st.release();
  if (e != cudaSuccess) return fail(g, TFHE_ERR_CUDA, "key upload: %s", cudaGetErrorString(e));
  if (rc && is_group(g) && g->err.empty()) g->err = c->err;
  return rc;
}

// cloudkey.NewCloudKey (cloudkey/cloudkey.go:24-145) on the device; see keygen.cuh.
int tfhe_ctx_generate_cloudkey(tfhe_ctx* g, const uint32_t* key_lv0, const uint32_t* key_lv1, double alpha_lv0,
                               double alpha_lv1, uint64_t seed, int with_ksk, uint32_t* offset_out, double* bsk_fft_out,
                               uint32_t* ksk_out, uint32_t* testvec_out) {
  if (!g || !key_lv0 || !key_lv1) return fail(g, TFHE_ERR_ARG, "null argument");
  if (!(alpha_lv0 >= 0.0) || !(alpha_lv1 >= 0.0)) return fail(g, TFHE_ERR_ARG, "noise parameters must be >= 0");
  if (ksk_out && !with_ksk) return fail(g, TFHE_ERR_ARG, "ksk_out given but with_ksk = 0");
  tfhe_ctx* c = is_group(g) ? g->kids[0] : g;  // a group generates on its first device and replicates by peer copy
  int rc = set_device(c);
  if (rc) return rc;
  const tfhe_params& P = c->P;
  const size_t bsk_bytes = (size_t)P.n * 2 * P.L * 2 * P.N * sizeof(double);
  const size_t ksk_rows = (size_t)P.N * P.iks_t * (1u << P.basebit);
  const size_t ksk_bytes = ksk_rows * (P.n + 1) * 4;
  const size_t tv_bytes = (size_t)2 * P.N * 4;
  // genDecompositionOffset (cloudkey.go:60-71) and genTestvec (:74-85: A = 0, B = F64ToTorus(0.125))
  uint32_t offset = 0;
  for (int i = 0; i < P.L; i++) offset += (uint32_t)(1u << (P.bgbit - 1)) * (uint32_t)(1u << (32 - (i + 1) * P.bgbit));
  std::vector<uint32_t> tv((size_t)2 * P.N, 0u);
  for (int i = 0; i < P.N; i++) tv[P.N + i] = 0x20000000u;
  RngKey rkey;
  if (!rng_make_key(seed, &rkey)) return fail(g, TFHE_ERR_STATE, "no entropy source (getrandom / /dev/urandom) for key generation");
  DevBuf s0, s1, sb, sk, st;
  cudaStream_t s = c->stream;
  cudaError_t e = s0.reserve((size_t)P.n * 4);
  if (e == cudaSuccess) e = s1.reserve((size_t)P.N * 4);
  if (e == cudaSuccess) e = sb.reserve(bsk_bytes);
  if (e == cudaSuccess && with_ksk) e = sk.reserve(ksk_bytes);
  if (e == cudaSuccess) e = st.reserve(tv_bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s0.p, key_lv0, (size_t)P.n * 4, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s1.p, key_lv1, (size_t)P.N * 4, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(st.p, tv.data(), tv_bytes, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    KeygenBskArgs a{};
    a.bsk_fft = sb.as<double>(); a.s0 = s0.as<uint32_t>(); a.s1 = s1.as<uint32_t>(); a.tw_tab = c->d_tw; a.alpha = alpha_lv1;
    a.key = rkey; a.L = P.L; a.bgbit = P.bgbit; a.tw0 = c->tw0;
    const unsigned grid = (unsigned)((size_t)P.n * 2 * P.L);
    const int T = P.N / 16;
    const size_t sm = (size_t)br_nbuf(c->logN) * TFHE_BR_EXW * (P.N / 2) * 16 + (size_t)P.N * 4 /* mask words */;
    switch (c->logN) {
      case 9: keygen_bsk_kernel<9><<<grid, T, sm, s>>>(a); break;
      case 10: keygen_bsk_kernel<10><<<grid, T, sm, s>>>(a); break;
      default: keygen_bsk_kernel<11><<<grid, T, sm, s>>>(a); break;
    }
    c->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && with_ksk) {
    keygen_ksk_kernel<<<(unsigned)ksk_rows, 128, 0, s>>>(sk.as<uint32_t>(), s0.as<uint32_t>(), s1.as<uint32_t>(), P.n, P.basebit,
                                                        P.iks_t, alpha_lv0, rkey);
    c->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    rc = tfhe_ctx_load_cloudkey_device(c, offset, sb.as<double>(), with_ksk ? sk.as<uint32_t>() : nullptr, st.as<uint32_t>(), s);
  if (e == cudaSuccess && rc == 0) {  // hand the CloudKey fields back in the reference's layouts
    if (offset_out) *offset_out = offset;
    if (bsk_fft_out) e = cudaMemcpy(bsk_fft_out, sb.p, bsk_bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && ksk_out) e = cudaMemcpy(ksk_out, sk.p, ksk_bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && testvec_out) memcpy(testvec_out, tv.data(), tv_bytes);
  }
  cudaStreamSynchronize(s);
  if (e == cudaSuccess && rc == 0 && is_group(g))
    rc = group_replicate(g, offset, sb.p, bsk_bytes, with_ksk ? sk.p : nullptr, ksk_bytes, st.p, tv_bytes);
  set_device(c);
  s0.release(); s1.release(); sb.release(); sk.release(); st.release();
  if (e != cudaSuccess) return fail(g, TFHE_ERR_CUDA, "key generation: %s", cudaGetErrorString(e));
  if (rc && is_group(g) && g->err.empty()) g->err = c->err;
  return rc;
}

// ---- device-buffer API ----------------------------------------------------------------------------
int tfhe_bootstrap_batch_device(tfhe_ctx* c, int64_t count, const uint32_t* d_ct_in, const uint32_t* d_luts,
                                int64_t nluts, uint32_t* d_ct_out, void* stream) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  if (is_group(c)) return fail(c, TFHE_ERR_ARG, "device-buffer entry points need a single-device context");
  if (count < 0 || (count > 0 && (!d_ct_in || !d_ct_out))) return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (d_luts && nluts != 1 && nluts != count) return fail(c, TFHE_ERR_ARG, "nluts must be 1 or count");
  if ((rc = set_device(c))) return rc;
  return bootstrap_device(c, count, d_ct_in, d_luts, nluts, d_ct_out, (cudaStream_t)stream);
}

}  // extern "C"

namespace {

// One batch of gates on device buffers, enqueued on `s`, never synchronised.  `ops` is HOST memory.
// Every gate that bootstraps becomes a JOB of a compacted batch: plain two-input gates first, then AND(a,b) of every MUX,
// then ANDNY(a,c) of every MUX (gates.go:107-114 with AND(NOT a, c) == ANDNY(a, c)); the second MUX level is
// OR(job, job).  Prepared rows are written straight into the compacted batch (never into d_out), NOT / COPY results
// straight into d_out, so d_out may be the very same buffer as d_a, d_b or d_c.  The index lists travel through the
// slot's pinned host buffer; `sl.idx_done` guards its reuse.
int gate_batch_device_impl(tfhe_ctx* c, int64_t count, const uint8_t* ops, int64_t nops, const uint32_t* d_a,
                           const uint32_t* d_b, const uint32_t* d_c, uint32_t* d_out, cudaStream_t s, Slot& sl) {
  const int n1 = c->P.n + 1;
  bool any_mux = false, any_unary = false, any_boot = false;
  for (int64_t g = 0; g < nops; g++) {
    const uint8_t op = ops[g];
    if (op > TFHE_OP_COPY) return fail(c, TFHE_ERR_ARG, "unknown opcode %d at gate %lld", (int)op, (long long)g);
    if (op == TFHE_OP_MUX) any_mux = true;
    else if (op >= TFHE_OP_NOT) any_unary = true;
    else any_boot = true;
  }
  if ((any_boot || any_mux) && !d_b) return fail(c, TFHE_ERR_ARG, "b is required by two-input gates");
  if (any_mux && !d_c) return fail(c, TFHE_ERR_ARG, "c is required by MUX gates");
  const bool uniform = nops == 1;
  // host staging: [job_of: count ints][src1 + muxg: <= count ints][opcodes: count bytes], pinned, one per slot
  const size_t need = (size_t)(uniform && !any_mux && !any_unary ? 0 : 2 * count) * sizeof(int) + (uniform ? 0 : (size_t)count);
  if (need) {
    if (sl.idx_pending) { CK(c, cudaEventSynchronize(sl.idx_done)); sl.idx_pending = false; }
    if (need > sl.h_idx_cap) {
      if (sl.h_idx) cudaFreeHost(sl.h_idx);
      sl.h_idx = nullptr; sl.h_idx_cap = 0;
      CK(c, cudaMallocHost(&sl.h_idx, need + need / 4));
      sl.h_idx_cap = need + need / 4;
    }
    CK(c, c->idx_a.reserve(need));
  }
  const uint8_t* d_ops = nullptr;
  if (!uniform) {  // opcodes: copied out of the caller's buffer before the call returns
    uint8_t* h = reinterpret_cast<uint8_t*>(sl.h_idx) + (size_t)(any_mux || any_unary ? 2 * count : 0) * sizeof(int);
    memcpy(h, ops, (size_t)count);
    d_ops = c->idx_a.as<uint8_t>() + (h - reinterpret_cast<uint8_t*>(sl.h_idx));
  }
  if (!any_mux && !any_unary) {  // fast path: job g = gate g, the key switch writes d_out directly
    if (!uniform) {
      CK(c, cudaMemcpyAsync(c->idx_a.p, sl.h_idx, need, cudaMemcpyHostToDevice, s));
      CK(c, cudaEventRecord(sl.idx_done, s));
      sl.idx_pending = true;
    }
    CK(c, c->prep.reserve((size_t)count * n1 * 4));
    gate_prepare_kernel<<<(unsigned)count, 256, 0, s>>>(count, d_ops, uniform ? (int)ops[0] : -1, d_a, d_b, d_c, nullptr,
                                                        c->prep.as<uint32_t>(), 0, 0, d_out, c->P.n);
    c->launches++;
    CK(c, cudaGetLastError());
    return bootstrap_device(c, count, c->prep.as<uint32_t>(), nullptr, 0, d_out, s);
  }
  int* job_of = sl.h_idx;
  int* lists = sl.h_idx + count;
  int64_t nb = 0, nm = 0;
  for (int64_t g = 0; g < count; g++) {
    const uint8_t op = ops[uniform ? 0 : g];
    if (op < TFHE_OP_MUX) nb++;
    else if (op == TFHE_OP_MUX) nm++;
  }
  int* src1 = lists;
  int* muxg = lists + nb;
  int64_t ib = 0, im = 0;
  for (int64_t g = 0; g < count; g++) {
    const uint8_t op = ops[uniform ? 0 : g];
    if (op < TFHE_OP_MUX) { job_of[g] = (int)ib; src1[ib++] = (int)g; }
    else if (op == TFHE_OP_MUX) { job_of[g] = (int)im; muxg[im++] = (int)g; }
    else job_of[g] = -1;
  }
  CK(c, cudaMemcpyAsync(c->idx_a.p, sl.h_idx, need, cudaMemcpyHostToDevice, s));
  CK(c, cudaEventRecord(sl.idx_done, s));
  sl.idx_pending = true;
  const int* d_job_of = c->idx_a.as<int>();
  const int* d_src1 = d_job_of + count;
  const int* d_muxg = d_src1 + nb;
  const int64_t j1 = nb + 2 * nm;
  CK(c, c->prep.reserve((size_t)std::max<int64_t>(j1, 1) * n1 * 4));
  CK(c, c->tmp.reserve((size_t)std::max<int64_t>(j1, 1) * n1 * 4));
  uint32_t* in1 = c->prep.as<uint32_t>();
  uint32_t* out1 = c->tmp.as<uint32_t>();
  gate_prepare_kernel<<<(unsigned)count, 256, 0, s>>>(count, d_ops, uniform ? (int)ops[0] : -1, d_a, d_b, d_c, d_job_of, in1,
                                                      (long long)nb, (long long)nm, d_out, c->P.n);
  c->launches++;
  CK(c, cudaGetLastError());
  int rc;
  if (c->mux_mode == 1 && nm) {  // opt-in: MUX = KeySwitch(BR(AND(a,b)) + BR(ANDNY(a,c)) + 1/8) — two blind rotations, one key switch
    const int N1 = c->P.N + 1;
    CK(c, c->lwe1.reserve((size_t)j1 * N1 * 4));
    CK(c, c->prep2.reserve((size_t)nm * N1 * 4));
    uint32_t* ext = c->lwe1.as<uint32_t>();
    if ((rc = launch_blind_rotate(c, j1, in1, nullptr, 0, ext, 1, s))) return rc;
    if (nb) {
      if ((rc = launch_key_switch(c, nb, ext, out1, s))) return rc;
      scatter_rows_kernel<<<(unsigned)nb, 128, 0, s>>>(out1, d_src1, d_out, n1);
      c->launches++;
    }
    mux_sum_kernel<<<(unsigned)nm, 256, 0, s>>>(ext + (size_t)nb * N1, ext + (size_t)(nb + nm) * N1, c->prep2.as<uint32_t>(), c->P.N);
    c->launches++;
    CK(c, cudaGetLastError());
    if ((rc = launch_key_switch(c, nm, c->prep2.as<uint32_t>(), out1, s))) return rc;
    scatter_rows_kernel<<<(unsigned)nm, 128, 0, s>>>(out1, d_muxg, d_out, n1);
    c->launches++;
    CK(c, cudaGetLastError());
    return TFHE_OK;
  }
  if (j1 && (rc = bootstrap_device(c, j1, in1, nullptr, 0, out1, s))) return rc;
  if (nb) {
    scatter_rows_kernel<<<(unsigned)nb, 128, 0, s>>>(out1, d_src1, d_out, n1);
    c->launches++;
  }
  if (nm) {  // level 2: OR(AND(a,b), ANDNY(a,c)); level-1 inputs are consumed by now (stream order), reuse their rows
    uint32_t* in2 = in1;
    mux_or_prepare_kernel<<<(unsigned)nm, 256, 0, s>>>(out1 + (size_t)nb * n1, out1 + (size_t)(nb + nm) * n1, in2, c->P.n);
    c->launches++;
    CK(c, cudaGetLastError());
    uint32_t* out2 = out1;
    if ((rc = bootstrap_device(c, nm, in2, nullptr, 0, out2, s))) return rc;
    scatter_rows_kernel<<<(unsigned)nm, 128, 0, s>>>(out2, d_muxg, d_out, n1);
    c->launches++;
  }
  CK(c, cudaGetLastError());
  return TFHE_OK;
}

// --- pipelined host-buffer calls -------------------------------------------------------------------------------------
// A host-buffer batch larger than ~1.5 pipeline chunks is cut into chunks that flow through two staging slots on three
// streams (copy in / compute / copy out), so that the transfers of chunk k+1 and k-1 overlap the compute of chunk k
// and device staging stays bounded whatever the batch size.  `in[]` are the per-row input arrays (a, b, c / ct, luts).
struct HostIn { const void* p; size_t row_bytes; DevBuf Slot::*buf; };

template <class Compute>  // Compute(slot, g0, cnt) enqueues the chunk's kernels on c->stream, reading slot.*, writing slot.out
int run_pipelined(tfhe_ctx* c, int64_t count, const HostIn* in, int nin, void* out, size_t out_row_bytes, Compute compute) {
  const int64_t CH = c->pipe_chunk;
  const int64_t nch = count <= CH + CH / 2 ? 1 : (count + CH - 1) / CH;
  const int64_t per = (count + nch - 1) / nch;
  for (int k = 0; k < (nch > 1 ? 2 : 1); k++) {
    Slot& sl = c->slot[k];
    for (int q = 0; q < nin; q++)
      if (in[q].p) CK(c, (sl.*(in[q].buf)).reserve((size_t)per * in[q].row_bytes));
    CK(c, sl.out.reserve((size_t)per * out_row_bytes));
  }
  cudaStream_t s_in = nch > 1 ? c->s_in : c->stream, s_out = nch > 1 ? c->s_out : c->stream;
  auto copy_out = [&](int64_t k) -> int {
    Slot& sl = c->slot[k & 1];
    const int64_t g0 = k * per, cnt = std::min<int64_t>(per, count - g0);
    if (nch > 1) CK(c, cudaStreamWaitEvent(s_out, sl.comp_done, 0));
    CK(c, cudaMemcpyAsync(reinterpret_cast<char*>(out) + (size_t)g0 * out_row_bytes, sl.out.p, (size_t)cnt * out_row_bytes,
                          cudaMemcpyDeviceToHost, s_out));
    if (nch > 1) CK(c, cudaEventRecord(sl.out_done, s_out));
    return 0;
  };
  for (int64_t k = 0; k < nch; k++) {
    Slot& sl = c->slot[k & 1];
    const int64_t g0 = k * per, cnt = std::min<int64_t>(per, count - g0);
    if (nch > 1 && k >= 2) CK(c, cudaStreamWaitEvent(s_in, sl.comp_done, 0));  // chunk k-2 has consumed this slot's inputs
    for (int q = 0; q < nin; q++)
      if (in[q].p)
        CK(c, cudaMemcpyAsync((sl.*(in[q].buf)).p, reinterpret_cast<const char*>(in[q].p) + (size_t)g0 * in[q].row_bytes,
                              (size_t)cnt * in[q].row_bytes, cudaMemcpyHostToDevice, s_in));
    if (nch > 1) {
      CK(c, cudaEventRecord(sl.in_done, s_in));
      CK(c, cudaStreamWaitEvent(c->stream, sl.in_done, 0));
      if (k >= 2) CK(c, cudaStreamWaitEvent(c->stream, sl.out_done, 0));  // chunk k-2's results have left this slot
    }
    int rc = compute(sl, g0, cnt);
    if (rc) return rc;
    if (nch > 1) CK(c, cudaEventRecord(sl.comp_done, c->stream));
    // results of the PREVIOUS chunk go out only now: with pageable host memory the copy blocks this thread until that
    // chunk is done, and the chunk just enqueued keeps the GPU busy meanwhile
    if (k >= 1 && (rc = copy_out(k - 1))) return rc;
  }
  int rc = copy_out(nch - 1);
  if (rc) return rc;
  CK(c, cudaStreamSynchronize(s_out));
  if (nch > 1) CK(c, cudaStreamSynchronize(c->stream));
  return TFHE_OK;
}

}  // namespace

extern "C" {

// ops is a HOST array (one byte per gate, or one for all); ciphertext pointers are device pointers.
int tfhe_gate_batch_device(tfhe_ctx* c, int64_t count, const uint8_t* ops, int64_t nops, const uint32_t* d_a,
                           const uint32_t* d_b, const uint32_t* d_c, uint32_t* d_out, void* stream) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  if (is_group(c)) return fail(c, TFHE_ERR_ARG, "device-buffer entry points need a single-device context");
  if (count < 0 || !ops || (nops != 1 && nops != count) || (count > 0 && (!d_a || !d_out)))
    return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (count == 0) return TFHE_OK;
  if ((rc = set_device(c))) return rc;
  return gate_batch_device_impl(c, count, ops, nops, d_a, d_b, d_c, d_out, (cudaStream_t)stream, c->slot[0]);
}

// ---- host-buffer API --------------------------------------------------------------------------------
static int h2d(tfhe_ctx* c, DevBuf& b, const void* src, size_t bytes) {
  CK(c, b.reserve(bytes));
  CK(c, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

// Bootstraps with host buffers: the reference's form (luts == NULL / one LUT / a LUT per ciphertext), a LUT table with an
// index per ciphertext (lut_index != NULL), and many-LUT bootstraps (log2k > 0: 2^log2k outputs per ciphertext).
static int bootstrap_host(tfhe_ctx* c, int64_t count, const uint32_t* ct_in, const uint32_t* luts, int64_t nluts,
                          const int32_t* lut_index, int log2k, uint32_t* ct_out) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  if (count < 0 || (count > 0 && (!ct_in || !ct_out))) return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (lut_index && (!luts || nluts < 1)) return fail(c, TFHE_ERR_ARG, "lut_index needs a LUT table");
  if (luts && !lut_index && nluts != 1 && nluts != count) return fail(c, TFHE_ERR_ARG, "nluts must be 1 or count");
  if (log2k < 0 || (1 << log2k) > c->P.N / 2) return fail(c, TFHE_ERR_ARG, "log2_k out of range");
  if (log2k > 0 && !luts) return fail(c, TFHE_ERR_ARG, "a many-LUT bootstrap needs a packed LUT");
  if (count == 0) return TFHE_OK;
  if (lut_index)
    for (int64_t g = 0; g < count; g++)
      if (lut_index[g] < 0 || lut_index[g] >= nluts) return fail(c, TFHE_ERR_ARG, "lut_index[%lld] = %d out of range", (long long)g, (int)lut_index[g]);
  const int64_t per = log2k > 0 ? (1ll << log2k) : 1;
  const size_t n1 = (size_t)c->P.n + 1, row = n1 * 4, lrow = (size_t)2 * c->P.N * 4;
  const bool per_ct = luts && !lut_index && nluts != 1;
  if (is_group(c)) {
    const std::vector<int64_t> bd = shard_bounds((int)c->kids.size(), count, nullptr, 0);
    return for_each_kid(c, [&](int d) {
      const int64_t g0 = bd[d], cnt = bd[d + 1] - bd[d];
      return bootstrap_host(c->kids[d], cnt, ct_in + (size_t)g0 * n1, luts ? luts + (per_ct ? (size_t)g0 * 2 * c->P.N : 0) : nullptr,
                            luts ? (per_ct ? cnt : nluts) : 0, lut_index ? lut_index + g0 : nullptr, log2k,
                            ct_out + (size_t)g0 * per * n1);
    });
  }
  if ((rc = set_device(c))) return rc;
  if (luts && !per_ct && (rc = h2d(c, c->h2d_luts, luts, (size_t)nluts * lrow))) return rc;
  const HostIn in[3] = {{ct_in, row, &Slot::a}, {per_ct ? luts : nullptr, lrow, &Slot::luts}, {lut_index, 4, &Slot::c}};
  return run_pipelined(c, count, in, 3, ct_out, (size_t)per * row, [&](Slot& sl, int64_t, int64_t cnt) {
    BrOpts opt;
    opt.d_lut_index = lut_index ? sl.c.as<int>() : nullptr;
    opt.ms_log2k = log2k;
    opt.extract_k = log2k > 0 ? (int)per : 0;
    return bootstrap_device(c, cnt, sl.a.as<uint32_t>(), luts ? (per_ct ? sl.luts.as<uint32_t>() : c->h2d_luts.as<uint32_t>()) : nullptr,
                            luts ? (per_ct ? cnt : nluts) : 0, sl.out.as<uint32_t>(), c->stream, nullptr, 1, opt);
  });
}

int tfhe_bootstrap_batch(tfhe_ctx* c, int64_t count, const uint32_t* ct_in, const uint32_t* luts, int64_t nluts,
                         uint32_t* ct_out) {
  if (!c) return TFHE_ERR_ARG;
  return bootstrap_host(c, count, ct_in, luts, nluts, nullptr, 0, ct_out);
}

int tfhe_bootstrap_batch_indexed(tfhe_ctx* c, int64_t count, const uint32_t* ct_in, const uint32_t* luts, int64_t nluts,
                                 const int32_t* lut_index, uint32_t* ct_out) {
  if (!c) return TFHE_ERR_ARG;
  if (!lut_index) return fail(c, TFHE_ERR_ARG, "lut_index is NULL");
  return bootstrap_host(c, count, ct_in, luts, nluts, lut_index, 0, ct_out);
}

int tfhe_bootstrap_multi_lut_batch(tfhe_ctx* c, int64_t count, const uint32_t* ct_in, const uint32_t* packed_luts,
                                   int64_t nluts, int32_t log2_k, uint32_t* ct_out) {
  if (!c) return TFHE_ERR_ARG;
  if (log2_k < 1) return fail(c, TFHE_ERR_ARG, "log2_k must be >= 1 (use tfhe_bootstrap_batch for one function)");
  return bootstrap_host(c, count, ct_in, packed_luts, nluts, nullptr, log2_k, ct_out);
}

int tfhe_gate_batch(tfhe_ctx* c, int64_t count, const uint8_t* ops, int64_t nops, const uint32_t* a, const uint32_t* b,
                    const uint32_t* cc, uint32_t* out) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  if (count < 0 || !ops || (nops != 1 && nops != count) || (count > 0 && (!a || !out)))
    return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (count == 0) return TFHE_OK;
  const size_t n1 = (size_t)c->P.n + 1;
  if (is_group(c)) {  // shards of equal bootstrap cost (MUX = 3), one host thread per device
    for (int64_t g = 0; g < nops; g++)
      if (ops[g] > TFHE_OP_COPY) return fail(c, TFHE_ERR_ARG, "unknown opcode %d at gate %lld", (int)ops[g], (long long)g);
    const std::vector<int64_t> bd = shard_bounds((int)c->kids.size(), count, ops, nops);
    return for_each_kid(c, [&](int d) {
      const int64_t g0 = bd[d], cnt = bd[d + 1] - bd[d];
      return tfhe_gate_batch(c->kids[d], cnt, ops + (nops == 1 ? 0 : g0), nops == 1 ? 1 : cnt, a + g0 * n1, b ? b + g0 * n1 : nullptr,
                             cc ? cc + g0 * n1 : nullptr, out + g0 * n1);
    });
  }
  if ((rc = set_device(c))) return rc;
  const HostIn in[3] = {{a, n1 * 4, &Slot::a}, {b, n1 * 4, &Slot::b}, {cc, n1 * 4, &Slot::c}};
  return run_pipelined(c, count, in, 3, out, n1 * 4, [&](Slot& sl, int64_t g0, int64_t cnt) {
    return gate_batch_device_impl(c, cnt, ops + (nops == 1 ? 0 : g0), nops == 1 ? 1 : cnt, sl.a.as<uint32_t>(),
                                  b ? sl.b.as<uint32_t>() : nullptr, cc ? sl.c.as<uint32_t>() : nullptr, sl.out.as<uint32_t>(),
                                  c->stream, sl);
  });
}

int tfhe_blind_rotate_batch(tfhe_ctx* c, int64_t count, const uint32_t* ct_in, const uint32_t* luts, int64_t nluts,
                            uint32_t* trlwe_out) {
  int rc = check_ready(c, false);
  if (rc) return rc;
  if (count < 0 || (count > 0 && (!ct_in || !trlwe_out))) return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (luts && nluts != 1 && nluts != count) return fail(c, TFHE_ERR_ARG, "nluts must be 1 or count");
  if (count == 0) return TFHE_OK;
  const size_t row = (size_t)(c->P.n + 1) * 4, lrow = (size_t)2 * c->P.N * 4;
  if (is_group(c)) {
    const std::vector<int64_t> b = shard_bounds((int)c->kids.size(), count, nullptr, 0);
    return for_each_kid(c, [&](int d) {
      const int64_t g0 = b[d], cnt = b[d + 1] - b[d];
      return tfhe_blind_rotate_batch(c->kids[d], cnt, ct_in + (size_t)g0 * (c->P.n + 1),
                                     luts ? luts + (nluts == 1 ? 0 : (size_t)g0 * 2 * c->P.N) : nullptr, luts ? (nluts == 1 ? 1 : cnt) : 0,
                                     trlwe_out + (size_t)g0 * 2 * c->P.N);
    });
  }
  if ((rc = set_device(c))) return rc;
  const bool per_ct = luts && nluts != 1;
  if (luts && !per_ct && (rc = h2d(c, c->h2d_luts, luts, lrow))) return rc;
  const HostIn in[2] = {{ct_in, row, &Slot::a}, {per_ct ? luts : nullptr, lrow, &Slot::luts}};
  return run_pipelined(c, count, in, 2, trlwe_out, lrow, [&](Slot& sl, int64_t, int64_t cnt) {
    return launch_blind_rotate(c, cnt, sl.a.as<uint32_t>(), luts ? (per_ct ? sl.luts.as<uint32_t>() : c->h2d_luts.as<uint32_t>()) : nullptr,
                               luts ? (per_ct ? cnt : 1) : 0, sl.out.as<uint32_t>(), 0, c->stream);
  });
}

int tfhe_cmux_batch(tfhe_ctx* c, int64_t count, int32_t bsk_index, const uint32_t* ct0, const uint32_t* ct1,
                    uint32_t* out) {
  int rc = check_ready(c, false);
  if (rc) return rc;
  if (is_group(c)) return tfhe_cmux_batch(c->kids[0], count, bsk_index, ct0, ct1, out) ? fail(c, TFHE_ERR_CUDA, "%s", c->kids[0]->err.c_str()) : TFHE_OK;
  if (count < 0 || bsk_index < 0 || bsk_index >= c->P.n || (count > 0 && (!ct1 || !out)))
    return fail(c, TFHE_ERR_ARG, "bad cmux arguments");
  if (count == 0) return TFHE_OK;
  if ((rc = set_device(c))) return rc;
  const size_t bytes = (size_t)count * 2 * c->P.N * 4;
  if (ct0 && (rc = h2d(c, c->h2d_a, ct0, bytes))) return rc;
  if ((rc = h2d(c, c->h2d_b, ct1, bytes))) return rc;
  CK(c, c->d2h_out.reserve(bytes));
  const Variant& V = kVariants[c->variant];
  CmuxArgs a{};
  a.ct0 = ct0 ? c->h2d_a.as<uint32_t>() : nullptr;
  a.ct1 = c->h2d_b.as<uint32_t>();
  a.bsk_row = c->d_bsk + (size_t)bsk_index * 2 * c->P.L * 2 * (c->P.N / 2);
  a.tw_tab = c->d_tw; a.out = c->d2h_out.as<uint32_t>(); a.offset = c->offset; a.tw0 = c->tw0;
  V.cmux<<<(unsigned)count, c->P.N / 16, cmux_smem(c->P.N), c->stream>>>(a);
  c->launches++;
  CK(c, cudaGetLastError());
  CK(c, cudaMemcpyAsync(out, c->d2h_out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return TFHE_OK;
}

int tfhe_sample_extract_batch(tfhe_ctx* c, int64_t count, const uint32_t* trlwe_in, uint32_t* lwe_out) {
  if (!c) return TFHE_ERR_ARG;
  if (is_group(c)) return tfhe_sample_extract_batch(c->kids[0], count, trlwe_in, lwe_out) ? fail(c, TFHE_ERR_CUDA, "%s", c->kids[0]->err.c_str()) : TFHE_OK;
  if (count < 0 || (count > 0 && (!trlwe_in || !lwe_out))) return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (count == 0) return TFHE_OK;
  int rc = set_device(c);
  if (rc) return rc;
  const size_t in_bytes = (size_t)count * 2 * c->P.N * 4, out_bytes = (size_t)count * (c->P.N + 1) * 4;
  if ((rc = h2d(c, c->h2d_a, trlwe_in, in_bytes))) return rc;
  CK(c, c->d2h_out.reserve(out_bytes));
  sample_extract_kernel<<<(unsigned)count, 256, 0, c->stream>>>(c->h2d_a.as<uint32_t>(), c->d2h_out.as<uint32_t>(), c->P.N);
  c->launches++;
  CK(c, cudaGetLastError());
  CK(c, cudaMemcpyAsync(lwe_out, c->d2h_out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return TFHE_OK;
}

int tfhe_key_switch_batch(tfhe_ctx* c, int64_t count, const uint32_t* lwe_in, uint32_t* ct_out) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  if (is_group(c)) return tfhe_key_switch_batch(c->kids[0], count, lwe_in, ct_out) ? fail(c, TFHE_ERR_CUDA, "%s", c->kids[0]->err.c_str()) : TFHE_OK;
  if (count < 0 || (count > 0 && (!lwe_in || !ct_out))) return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (count == 0) return TFHE_OK;
  if ((rc = set_device(c))) return rc;
  const size_t in_bytes = (size_t)count * (c->P.N + 1) * 4, out_bytes = (size_t)count * (c->P.n + 1) * 4;
  if ((rc = h2d(c, c->h2d_a, lwe_in, in_bytes))) return rc;
  CK(c, c->d2h_out.reserve(out_bytes));
  if ((rc = launch_key_switch(c, count, c->h2d_a.as<uint32_t>(), c->d2h_out.as<uint32_t>(), c->stream))) return rc;
  CK(c, cudaMemcpyAsync(ct_out, c->d2h_out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return TFHE_OK;
}


static int h2d(tfhe_ctx* c, DevBuf& b, const void* src, size_t bytes);

// ---- proxy re-encryption on the key-switch kernel (proxyreenc/proxyreenc.go:321-366) ------------------------------------
// ReencryptTLWELv0 is IdentityKeySwitching with source dimension n instead of N: out = (0,..,0,b) - sum over the non-zero
// digits k of a_i + 2^(31 - basebit t) of KeyEncryptions[base t i + base j + k].  The key is uploaded once per context.
int tfhe_ctx_load_reencryption_key(tfhe_ctx* c, const uint32_t* key_encryptions, int32_t basebit, int32_t t) {
  GROUP_FORWARD(c, tfhe_ctx_load_reencryption_key(kid, key_encryptions, basebit, t));
  if (!c || !key_encryptions) return fail(c, TFHE_ERR_ARG, "null argument");
  if (basebit < 1 || t < 1 || basebit * t > 31) return fail(c, TFHE_ERR_ARG, "invalid re-encryption decomposition");
  int rc = set_device(c);
  if (rc) return rc;
  const int n = c->P.n;
  if ((size_t)n * t * 4 > 128 * 1024) return fail(c, TFHE_ERR_ARG, "n * t too large for the key-switch kernel");
  const size_t rows = (size_t)n * t * (1u << basebit);
  const int stride = (n + 1 + 3) / 4 * 4;
  DevBuf st;
  CK(c, st.reserve(rows * (n + 1) * 4));
  if (c->d_reenc) { cudaFree(c->d_reenc); c->d_reenc = nullptr; }
  CK(c, cudaMalloc(&c->d_reenc, rows * stride * 4));
  CK(c, cudaMemcpyAsync(st.p, key_encryptions, rows * (n + 1) * 4, cudaMemcpyHostToDevice, c->stream));
  ksk_repack_kernel<<<(unsigned)rows, 128, 0, c->stream>>>(st.as<uint32_t>(), c->d_reenc, n + 1, stride, 1 << basebit);
  c->launches++;
  CK(c, cudaGetLastError());
  CK(c, cudaStreamSynchronize(c->stream));
  st.release();
  c->reenc_stride = stride; c->reenc_basebit = basebit; c->reenc_t = t;
  return TFHE_OK;
}

int tfhe_reencrypt_batch(tfhe_ctx* c, int64_t count, const uint32_t* ct_in, uint32_t* ct_out) {
  if (!c) return TFHE_ERR_ARG;
  if (count < 0 || (count > 0 && (!ct_in || !ct_out))) return fail(c, TFHE_ERR_ARG, "bad batch arguments");
  if (count == 0) return TFHE_OK;
  const size_t n1 = (size_t)c->P.n + 1;
  if (is_group(c)) {
    const std::vector<int64_t> bd = shard_bounds((int)c->kids.size(), count, nullptr, 0);
    return for_each_kid(c, [&](int d) {
      return tfhe_reencrypt_batch(c->kids[d], bd[d + 1] - bd[d], ct_in + (size_t)bd[d] * n1, ct_out + (size_t)bd[d] * n1);
    });
  }
  if (!c->d_reenc) return fail(c, TFHE_ERR_STATE, "re-encryption key not loaded");
  int rc = set_device(c);
  if (rc) return rc;
  const int n = c->P.n;
  const size_t sm = (size_t)n * c->reenc_t * sizeof(uint32_t);
  const int threads = std::min(512, std::max(128, (c->reenc_stride / 4 + 31) / 32 * 32));
  const HostIn in[1] = {{ct_in, n1 * 4, &Slot::a}};
  return run_pipelined(c, count, in, 1, ct_out, n1 * 4, [&](Slot& sl, int64_t, int64_t cnt) {
    const int splits = (int)std::max<int64_t>(1, std::min<int64_t>(64, (2 * (int64_t)c->sm_count) / cnt));
    if (splits > 1) {
      zero_out_rows_kernel<<<(unsigned)cnt, 128, 0, c->stream>>>(sl.out.as<uint32_t>(), n, nullptr, 1);
      c->launches++;
    }
    key_switch_kernel<<<dim3((unsigned)cnt, (unsigned)splits), threads, sm, c->stream>>>(
        sl.a.as<uint32_t>(), c->d_reenc, sl.out.as<uint32_t>(), /*source dimension*/ n, /*target dimension*/ n, c->reenc_basebit,
        c->reenc_t, c->reenc_stride, nullptr, 1, splits);
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
  });
}

// ---- levelised circuit runner --------------------------------------------------------------------------
int tfhe_circuit_run(tfhe_ctx* c, int64_t instances, int32_t n_inputs, int32_t n_gates, const tfhe_gate_desc* gates,
                     const uint32_t* inputs, int32_t n_outputs, const int32_t* output_wires, uint32_t* outputs) {
  int rc = check_ready(c, true);
  if (rc) return rc;
  if (instances < 0 || n_inputs < 0 || n_gates < 0 || n_outputs < 0 || (n_gates > 0 && !gates) ||
      (n_inputs > 0 && instances > 0 && !inputs) || (n_outputs > 0 && (!output_wires || (instances > 0 && !outputs))))
    return fail(c, TFHE_ERR_ARG, "bad circuit arguments");
  if (is_group(c)) {  // whole instances per device: no ciphertext ever crosses GPUs (wires are [wire][instance][n+1])
    const int nd = (int)c->kids.size();
    const size_t n1g = (size_t)c->P.n + 1;
    std::vector<std::vector<uint32_t>> in_sh(nd), out_sh(nd);
    std::vector<int64_t> bnd(nd + 1);
    for (int d = 0; d <= nd; d++) bnd[d] = instances * d / nd;
    rc = for_each_kid(c, [&](int d) -> int {
      const int64_t i0 = bnd[d], cnt = bnd[d + 1] - bnd[d];
      in_sh[d].resize((size_t)n_inputs * cnt * n1g);
      for (int w = 0; w < n_inputs; w++)
        if (cnt) memcpy(in_sh[d].data() + (size_t)w * cnt * n1g, inputs + ((size_t)w * instances + i0) * n1g, (size_t)cnt * n1g * 4);
      out_sh[d].resize((size_t)n_outputs * cnt * n1g);
      const int r = tfhe_circuit_run(c->kids[d], cnt, n_inputs, n_gates, gates, in_sh[d].data(), n_outputs, output_wires, out_sh[d].data());
      if (r) return r;
      for (int k = 0; k < n_outputs; k++)
        if (cnt) memcpy(outputs + ((size_t)k * instances + i0) * n1g, out_sh[d].data() + (size_t)k * cnt * n1g, (size_t)cnt * n1g * 4);
      return 0;
    });
    return rc;
  }
  if ((rc = set_device(c))) return rc;
  // 1. validate (topological order, single assignment), expand MUX, compute depths
  int n_wires = n_inputs;
  for (int g = 0; g < n_gates; g++) n_wires = gates[g].out + 1 > n_wires ? gates[g].out + 1 : n_wires;
  std::vector<int> depth(n_wires, -1);
  for (int w = 0; w < n_inputs; w++) depth[w] = 0;
  struct HGate { GateDesc d; int depth; bool linear; };
  std::vector<HGate> hg;
  auto ready = [&](int w) { return w >= 0 && w < (int)depth.size() && depth[w] >= 0; };
  for (int g = 0; g < n_gates; g++) {
    const tfhe_gate_desc& q = gates[g];
    if (q.op > TFHE_OP_COPY) return fail(c, TFHE_ERR_ARG, "gate %d: unknown opcode %d", g, (int)q.op);
    const bool unary = q.op >= TFHE_OP_NOT, mux = q.op == TFHE_OP_MUX;
    if (!ready(q.in0) || (!unary && !ready(q.in1)) || (mux && !ready(q.in2)))
      return fail(c, TFHE_ERR_ARG, "gate %d reads a wire that is not an input or an earlier gate's output", g);
    if (q.out < n_inputs || q.out >= n_wires || depth[q.out] >= 0)
      return fail(c, TFHE_ERR_ARG, "gate %d: output wire %d is an input or already assigned", g, q.out);
    if (unary) {
      hg.push_back({{q.op, q.in0, q.in0, q.out}, depth[q.in0], true});
      depth[q.out] = depth[q.in0];
    } else if (mux) {  // OR(AND(a,b), ANDNY(a,c)) — gates.go:107-114 with AND(NOT a, c) == ANDNY(a, c)
      const int t0 = (int)depth.size(), t1 = t0 + 1;
      depth.push_back(-1); depth.push_back(-1);
      const int d1 = 1 + std::max(depth[q.in0], std::max(depth[q.in1], depth[q.in2]));
      hg.push_back({{TFHE_OP_AND, q.in0, q.in1, t0}, d1, false});
      hg.push_back({{TFHE_OP_ANDNY, q.in0, q.in2, t1}, d1, false});
      hg.push_back({{TFHE_OP_OR, t0, t1, q.out}, d1 + 1, false});
      depth[t0] = depth[t1] = d1;
      depth[q.out] = d1 + 1;
    } else {
      const int d1 = 1 + std::max(depth[q.in0], depth[q.in1]);
      hg.push_back({{q.op, q.in0, q.in1, q.out}, d1, false});
      depth[q.out] = d1;
    }
  }
  for (int k = 0; k < n_outputs; k++)
    if (!ready(output_wires[k])) return fail(c, TFHE_ERR_ARG, "output %d names an unassigned wire", k);
  if (instances == 0) return TFHE_OK;
  const int total_wires = (int)depth.size();
  int max_depth = 0;
  for (auto& h : hg) max_depth = std::max(max_depth, h.depth);
  // 2. device state
  const int n1 = c->P.n + 1;
  const size_t wire_bytes = (size_t)instances * n1 * 4;
  CK(c, c->wires.reserve((size_t)total_wires * wire_bytes));
  cudaStream_t s = c->stream;
  if (n_inputs) CK(c, cudaMemcpyAsync(c->wires.p, inputs, (size_t)n_inputs * wire_bytes, cudaMemcpyHostToDevice, s));
  std::vector<GateDesc> sorted;  // bootstrapped gates grouped by depth
  std::vector<int> level_off(max_depth + 2, 0);
  size_t widest = 0;
  for (int d = 1; d <= max_depth; d++) {
    level_off[d] = (int)sorted.size();
    for (auto& h : hg) if (!h.linear && h.depth == d) sorted.push_back(h.d);
    widest = std::max(widest, sorted.size() - (size_t)level_off[d]);
  }
  level_off[max_depth + 1] = (int)sorted.size();
  if (!sorted.empty()) {
    CK(c, c->gate_descs.reserve(sorted.size() * sizeof(GateDesc)));
    CK(c, cudaMemcpyAsync(c->gate_descs.p, sorted.data(), sorted.size() * sizeof(GateDesc), cudaMemcpyHostToDevice, s));
    CK(c, c->prep.reserve(widest * wire_bytes));
  }
  // 3. run: depth d = linear gates whose input has depth d (in list order), then the bootstrapped gates of depth d+1
  auto run_linear = [&](int d) -> int {
    for (auto& h : hg)
      if (h.linear && h.depth == d) {
        circuit_linear_kernel<<<256, 256, 0, s>>>(h.d, instances, c->wires.as<uint32_t>(), c->P.n);
        c->launches++;
      }
    CK(c, cudaGetLastError());
    return 0;
  };
  auto run_levels = [&]() -> int {
    int r = run_linear(0);
    if (r) return r;
    for (int d = 1; d <= max_depth; d++) {
      const int ng = level_off[d + 1] - level_off[d];
      if (ng > 0) {
        const GateDesc* dg = c->gate_descs.as<GateDesc>() + level_off[d];
        const int64_t jobs = (int64_t)ng * instances;
        circuit_prepare_kernel<<<(unsigned)jobs, 256, 0, s>>>(dg, instances, c->wires.as<uint32_t>(), c->prep.as<uint32_t>(), c->P.n);
        c->launches++;
        CK(c, cudaGetLastError());
        if ((r = bootstrap_device(c, jobs, c->prep.as<uint32_t>(), nullptr, 0, c->wires.as<uint32_t>(), s, dg, instances))) return r;
      }
      if ((r = run_linear(d))) return r;
    }
    return 0;
  };
  // CUDA-graph replay of the level loop.  First call with a given (gate list, inputs, instances): eager, which also sizes
  // every scratch buffer.  Second call: the same launches are captured into a graph (and run).  From then on the graph is
  // replayed — one launch for the whole circuit — as long as no device buffer has moved since the capture.
  bool done = false;
  if (c->circuit_graph && !c->timing) {
    tfhe_ctx::CircuitGraph* e = nullptr;
    for (auto& q : c->circuit_graphs)
      if (q.instances == instances && q.n_inputs == n_inputs && q.gates.size() == (size_t)n_gates &&
          (n_gates == 0 || !memcmp(q.gates.data(), gates, (size_t)n_gates * sizeof(tfhe_gate_desc)))) { e = &q; break; }
    const uint64_t gen = g_alloc_generation.load(std::memory_order_relaxed);
    if (e && e->exec && e->gen == gen) {
      CK(c, cudaGraphLaunch(e->exec, s));
      c->launches += e->launches;
      c->circuit_graph_replays++;
      done = true;
    } else if (e) {  // seen before (buffers are sized), or captured before a buffer moved: capture now
      if (e->exec) { cudaGraphExecDestroy(e->exec); e->exec = nullptr; }
      const int64_t l0 = c->launches;
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int r = run_levels();
        const cudaError_t ce = cudaStreamEndCapture(s, &graph);
        const bool moved = g_alloc_generation.load(std::memory_order_relaxed) != gen;
        if (r == 0 && ce == cudaSuccess && graph && !moved && cudaGraphInstantiate(&e->exec, graph, 0) == cudaSuccess) {
          e->gen = gen;
          e->launches = c->launches - l0;
          CK(c, cudaGraphLaunch(e->exec, s));
          done = true;
        } else {
          e->exec = nullptr;
          cudaGetLastError();    // a capture that could not be completed is not an error of the call: run eagerly below
          c->launches = l0;
        }
        if (graph) cudaGraphDestroy(graph);
      } else {
        cudaGetLastError();
      }
    } else {
      if (c->circuit_graphs.size() >= 8) {  // small cache, oldest entry out
        if (c->circuit_graphs.front().exec) cudaGraphExecDestroy(c->circuit_graphs.front().exec);
        c->circuit_graphs.erase(c->circuit_graphs.begin());
      }
      tfhe_ctx::CircuitGraph q;
      q.gates.assign(gates, gates + n_gates);
      q.n_inputs = n_inputs;
      q.instances = instances;
      c->circuit_graphs.push_back(std::move(q));
    }
  }
  if (!done && (rc = run_levels())) return rc;
  for (int k = 0; k < n_outputs; k++)
    CK(c, cudaMemcpyAsync(outputs + (size_t)k * instances * n1, c->wires.as<char>() + (size_t)output_wires[k] * wire_bytes, wire_bytes,
                          cudaMemcpyDeviceToHost, s));
  CK(c, cudaStreamSynchronize(s));
  return TFHE_OK;
}

// ---- stand-alone polynomial transforms (reference FourierPoly layout) --------------------------------------
static int poly_call(tfhe_ctx* c, int mode, int64_t count, const void* in0, size_t in0_bytes, const void* in1,
                     size_t in1_bytes, void* out, size_t out_bytes) {
  if (!c) return TFHE_ERR_ARG;
  if (is_group(c)) c = c->kids[0];
  if (count < 0 || (count > 0 && (!in0 || !out || (mode == 2 && !in1)))) return fail(c, TFHE_ERR_ARG, "bad polynomial batch arguments");
  if (count == 0) return TFHE_OK;
  int rc = set_device(c);
  if (rc) return rc;
  if ((rc = h2d(c, c->h2d_a, in0, in0_bytes))) return rc;
  if (mode == 2 && (rc = h2d(c, c->h2d_b, in1, in1_bytes))) return rc;
  CK(c, c->d2h_out.reserve(out_bytes));
  PolyArgs a{};
  a.in0 = c->h2d_a.p; a.in1 = c->h2d_b.p; a.out = c->d2h_out.p; a.tw_tab = c->d_tw; a.mode = mode; a.tw0 = c->tw0;
  const int N = c->P.N, T = N / 16;
  const size_t sm = (size_t)br_nbuf(c->logN) * TFHE_BR_EXW * (N / 2) * 16;
  switch (c->logN) {
    case 9: poly_kernel<9><<<(unsigned)count, T, sm, c->stream>>>(a); break;
    case 10: poly_kernel<10><<<(unsigned)count, T, sm, c->stream>>>(a); break;
    default: poly_kernel<11><<<(unsigned)count, T, sm, c->stream>>>(a); break;
  }
  c->launches++;
  CK(c, cudaGetLastError());
  CK(c, cudaMemcpyAsync(out, c->d2h_out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return TFHE_OK;
}

int tfhe_to_fourier_batch(tfhe_ctx* c, int64_t count, const uint32_t* poly_in, double* fourier_out) {
  const size_t N = c ? (size_t)c->P.N : 0;
  return poly_call(c, 0, count, poly_in, (size_t)count * N * 4, nullptr, 0, fourier_out, (size_t)count * N * 8);
}
int tfhe_to_poly_batch(tfhe_ctx* c, int64_t count, const double* fourier_in, uint32_t* poly_out) {
  const size_t N = c ? (size_t)c->P.N : 0;
  return poly_call(c, 1, count, fourier_in, (size_t)count * N * 8, nullptr, 0, poly_out, (size_t)count * N * 4);
}
int tfhe_mul_poly_batch(tfhe_ctx* c, int64_t count, const uint32_t* p0, const uint32_t* p1, uint32_t* out) {
  const size_t N = c ? (size_t)c->P.N : 0;
  return poly_call(c, 2, count, p0, (size_t)count * N * 4, p1, (size_t)count * N * 4, out, (size_t)count * N * 4);
}

int64_t tfhe_ctx_kernel_launches(const tfhe_ctx* c) {
  if (!c) return 0;
  int64_t total = c->launches;
  for (const tfhe_ctx* k : c->kids) total += k->launches;
  return total;
}

int tfhe_ctx_set_key_switch_variant(tfhe_ctx* c, int variant) {
  GROUP_FORWARD(c, tfhe_ctx_set_key_switch_variant(kid, variant));
  if (!c || variant < 0 || variant > 3) return fail(c, TFHE_ERR_ARG, "key-switch variant must be 0 (auto), 1 (gather), 2 (tensor core) or 3 (shared-memory tiles)");
  c->ks_variant = variant;
  return TFHE_OK;
}

int tfhe_ctx_set_blind_rotate_variant(tfhe_ctx* c, int variant) {
  GROUP_FORWARD(c, tfhe_ctx_set_blind_rotate_variant(kid, variant));
  if (!c) return TFHE_ERR_ARG;
  if (variant == 0) { c->br_variant = 0; c->br_auto_lat = true; return TFHE_OK; }
  if (variant == 10) { c->br_variant = 0; c->br_auto_lat = false; return TFHE_OK; }  // throughput kernel at every batch size
  if (variant == 12) {
    if (!kVariants[c->variant].br_latp) return fail(c, TFHE_ERR_ARG, "the order-preserving latency kernel exists for L <= 2 only");
    c->br_variant = 12; c->br_auto_lat = true; return TFHE_OK;
  }
  if (variant == 9) {
    if (!kVariants[c->variant].br_lat) return fail(c, TFHE_ERR_ARG, "the latency kernel exists for the exact N = 1024 sets only");
    c->br_variant = 9; c->br_auto_lat = true; return TFHE_OK;
  }
#if TFHE_EXPERIMENTAL
  if (variant == 13) {
    if (!kVariants[c->variant].br_cl) return fail(c, TFHE_ERR_ARG, "the cluster kernel exists for the exact N = 1024 sets only");
    c->br_variant = 13; c->br_auto_lat = true; return TFHE_OK;
  }
  if (variant == 11) {
    if (!kVariants[c->variant].br_lat2 || c->P.n > 2048) return fail(c, TFHE_ERR_ARG, "the per-digit latency kernel exists for the exact N = 1024 sets only");
    c->br_variant = 11; c->br_auto_lat = true; return TFHE_OK;
  }
  if (variant < 0 || variant > 8) return fail(c, TFHE_ERR_ARG, "unknown blind-rotate variant %d", variant);
  c->br_auto_lat = true;
  if (variant == 8 && (!kVariants[c->variant].br_mg || c->P.n > 2048)) return fail(c, TFHE_ERR_ARG, "the gates-per-block kernel exists for N = 1024, n <= 2048 only");
  if (variant == 7 && !kVariants[c->variant].br_tms) return fail(c, TFHE_ERR_ARG, "the six-blocks-per-SM kernel exists for N = 1024 only");
  if (variant >= 5 && !kVariants[c->variant].br_tx) return fail(c, TFHE_ERR_ARG, "the TMEM-exchange kernel exists for N = 1024 only");
  if (variant == 4 && !kVariants[c->variant].br_tm) return fail(c, TFHE_ERR_ARG, "the TMEM-accumulator kernel needs N >= 1024");
  if (variant == 3 && !kVariants[c->variant].br_w16) return fail(c, TFHE_ERR_ARG, "the warp-per-gate kernel exists for N = 1024 only");
  if (variant == 2 && c->key_loaded && !c->bsk_tex) return fail(c, TFHE_ERR_STATE, "no texture object");
  c->br_variant = variant;
  return TFHE_OK;
#else
  return fail(c, TFHE_ERR_ARG, "blind-rotate variant %d is an experimental kernel: rebuild with -DTFHE_EXPERIMENTAL=1 "
                               "(default library: 0 = automatic, 9 = lat, 10 = throughput kernel only, 12 = latp)", variant);
#endif
}

// MUX evaluation: 0 = gates.MUX of the reference (gates/gates.go:107-114: OR(AND(a,b), AND(NOT a, c)), three bootstraps,
// bit-identical to it), 1 = the two ANDs are blind-rotated and extracted WITHOUT key switch (gates.go:145-149), summed
// with 1/8 and key-switched once: one blind rotation fewer, same truth table, different (valid) ciphertext words.
int tfhe_ctx_set_mux_mode(tfhe_ctx* c, int mode) {
  GROUP_FORWARD(c, tfhe_ctx_set_mux_mode(kid, mode));
  if (!c || mode < 0 || mode > 1) return fail(c, TFHE_ERR_ARG, "mux mode must be 0 (three bootstraps) or 1 (two blind rotations + one key switch)");
  c->mux_mode = mode;
  return TFHE_OK;
}

// tfhe_circuit_run: 1 = capture the level loop of a repeated circuit into a CUDA graph and replay it (results identical).
int tfhe_ctx_set_circuit_graph(tfhe_ctx* c, int enable) {
  GROUP_FORWARD(c, tfhe_ctx_set_circuit_graph(kid, enable));
  if (!c) return TFHE_ERR_ARG;
  c->circuit_graph = enable ? 1 : 0;
  if (!enable) {
    for (auto& q : c->circuit_graphs)
      if (q.exec) cudaGraphExecDestroy(q.exec);
    c->circuit_graphs.clear();
  }
  return TFHE_OK;
}
int64_t tfhe_ctx_circuit_graph_replays(const tfhe_ctx* c) {
  if (!c) return 0;
  int64_t total = c->circuit_graph_replays;
  for (const tfhe_ctx* k : c->kids) total += k->circuit_graph_replays;
  return total;
}

// CMUX steps per work item of the persistent throughput kernel: 0 = automatic, >= n = whole gates per item.
int tfhe_ctx_set_blind_rotate_chunk_steps(tfhe_ctx* c, int steps) {
  GROUP_FORWARD(c, tfhe_ctx_set_blind_rotate_chunk_steps(kid, steps));
  if (!c || steps < 0) return fail(c, TFHE_ERR_ARG, "chunk steps must be >= 0");
  c->br_chunk_steps = steps;
  return TFHE_OK;
}

// ciphertexts per chunk of the pipelined host-buffer calls (default 16384; batches up to 1.5 chunks run unchunked)
int tfhe_ctx_set_pipeline_chunk(tfhe_ctx* c, int64_t rows) {
  GROUP_FORWARD(c, tfhe_ctx_set_pipeline_chunk(kid, rows));
  if (!c || rows < 1) return fail(c, TFHE_ERR_ARG, "pipeline chunk must be >= 1");
  c->pipe_chunk = rows;
  return TFHE_OK;
}

int tfhe_ctx_set_timing(tfhe_ctx* c, int enable) {
  GROUP_FORWARD(c, tfhe_ctx_set_timing(kid, enable));
  if (!c) return TFHE_ERR_ARG;
  c->timing = enable != 0;
  return TFHE_OK;
}

int tfhe_ctx_collect_timing(tfhe_ctx* c, double out[4]) {
  if (!c || !out) return TFHE_ERR_ARG;
  if (is_group(c)) {  // sums over the devices (they run concurrently: divide by tfhe_ctx_device_count for a per-device mean)
    out[0] = out[1] = out[2] = out[3] = 0.0;
    for (tfhe_ctx* k : c->kids) {
      double t[4];
      const int r = tfhe_ctx_collect_timing(k, t);
      if (r) { c->err = k->err; return r; }
      for (int q = 0; q < 4; q++) out[q] += t[q];
    }
    return TFHE_OK;
  }
  int rc = set_device(c);
  if (rc) return rc;
  out[0] = out[1] = out[2] = out[3] = 0.0;
  for (auto& ev : c->ev_live) {
    CK(c, cudaEventSynchronize(ev.e2));
    float a = 0.f, b = 0.f;
    CK(c, cudaEventElapsedTime(&a, ev.e0, ev.e1));
    CK(c, cudaEventElapsedTime(&b, ev.e1, ev.e2));
    out[0] += a; out[1] += 1.0; out[2] += b; out[3] += 1.0;
    c->ev_free.push_back(ev);
  }
  c->ev_live.clear();
  return TFHE_OK;
}

// Measured FP64 throughput of `device`: out = {TFLOP/s (2 flops per DFMA), thread-DFMA per clock per SM at the
// driver-reported SM clock, that clock in MHz}.  ~0.1 s.
int tfhe_fp64_peak_probe(int device, double out[3]) {
  if (!out) return TFHE_ERR_ARG;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, TFHE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  double *d = nullptr, *in = nullptr;
  if ((e = cudaMalloc(&d, 8)) != cudaSuccess || (e = cudaMalloc(&in, 8 * 32)) != cudaSuccess) {
    if (d) cudaFree(d);
    return fail(nullptr, TFHE_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  const double h[2] = {0.999999, 1e-9};
  cudaMemset(in, 0, 8 * 32);
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = sms * 8, iters = 1 << 15;  // 8 x 256 threads per SM = 16 warps per scheduler
  fp64_probe_kernel<<<blocks, 256>>>(d, in, 64);
  double best = 0.0;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    fp64_probe_kernel<<<blocks, 256>>>(d, in, iters);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)iters * 64.0 * 256.0 * blocks;
    if (ms > 0.f) best = std::max(best, dfma / (ms * 1e-3));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d); cudaFree(in);
  if (e != cudaSuccess) return fail(nullptr, TFHE_ERR_CUDA, "fp64 probe: %s", cudaGetErrorString(e));
  out[0] = 2.0 * best / 1e12;
  out[1] = khz > 0 ? best / ((double)khz * 1e3) / sms : 0.0;
  out[2] = khz / 1e3;
  return TFHE_OK;
}

int64_t tfhe_ctx_algorithmic_bytes_per_bootstrap(const tfhe_ctx* c) {
  if (!c) return 0;
  const tfhe_params& P = c->P;
  const int64_t base = 1ll << P.basebit;
  const int64_t bk = (int64_t)P.n * 2 * P.L * 2 * P.N * 8;
  const int64_t ks = (int64_t)P.N * P.iks_t * (base - 1) / base * (P.n + 1) * 4;
  const int64_t io = (int64_t)3 * (P.n + 1) * 4;
  return bk + ks + io;
}

}  // extern "C"
