// lwe_kernels.cuh — the integer (u32) kernels either side of the blind rotation:
//   gate prologues      evaluator/gates_helper.go:10-63, gates/gates.go:52-130
//   identity key switch trgsw/keyswitch.go:10-37 (= trgsw/trgsw.go:285-311)
//   sample extract      trlwe/trlwe_ops.go:10-21
//   key repacking       (layout only; no reference equivalent)
// All arithmetic is mod 2^32 and bit-exact with the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe {

// c = sa*a + sb*b + (0,...,0,bias) for one gate per block row.  (sa, sb, bias) per opcode:
//   NAND (-1,-1,+1/8)  AND (1,1,-1/8)   OR (1,1,+1/8)    XOR (1,2,+1/4)   XNOR (1,-2,+1/4)
//   NOR (-1,-1,-1/8)   ANDNY (-1,1,-1/8) ANDYN (1,-1,-1/8) ORNY (-1,1,+1/8) ORYN (1,-1,+1/8)
// with 1/8 = 0x20000000, 1/4 = 0x40000000, -1/8 = 0xE0000000 (utils/utils_test.go:15-19).
__device__ __forceinline__ void gate_coeffs(int op, uint32_t& sa, uint32_t& sb, uint32_t& bias) {
  const uint32_t P1 = 1u, M1 = 0xFFFFFFFFu, E8 = 0x20000000u, N8 = 0xE0000000u, Q4 = 0x40000000u;
  switch (op) {
    case 0: sa = M1; sb = M1; bias = E8; break;               // NAND
    case 1: sa = P1; sb = P1; bias = N8; break;               // AND
    case 2: sa = P1; sb = P1; bias = E8; break;               // OR
    case 3: sa = P1; sb = 2u; bias = Q4; break;               // XOR
    case 4: sa = P1; sb = 0xFFFFFFFEu; bias = Q4; break;      // XNOR
    case 5: sa = M1; sb = M1; bias = N8; break;               // NOR
    case 6: sa = M1; sb = P1; bias = N8; break;               // ANDNY
    case 7: sa = P1; sb = M1; bias = N8; break;               // ANDYN
    case 8: sa = M1; sb = P1; bias = E8; break;               // ORNY
    case 9: sa = P1; sb = M1; bias = E8; break;               // ORYN
    case 11: sa = M1; sb = 0u; bias = 0u; break;              // NOT  (0 - a, gates.go:117-119)
    default: sa = P1; sb = 0u; bias = 0u; break;              // COPY (gates.go:122-126)
  }
}

// grid.x = count, block = 256.  op_uniform >= 0: every gate has that opcode; else ops[g].  A gate that bootstraps
// writes its prepared ciphertext(s) into the compacted JOB batch `jobs`: job = job_of ? job_of[g] : g;
//   two-input gate  ->  jobs[job]                                   (job in [0, nb))
//   MUX (op 10)     ->  AND(a,b) at jobs[nb + job], ANDNY(a,c) at jobs[nb + nm + job]   (job in [0, nm);
//                       AND(NOT a, c) == ANDNY(a,c) word for word: (0 - a) + c - 1/8)
// NOT / COPY (no bootstrap) write their final result to direct_out[g].
__global__ void gate_prepare_kernel(long long count, const uint8_t* __restrict__ ops, int op_uniform,
                                    const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                    const uint32_t* __restrict__ c, const int* __restrict__ job_of,
                                    uint32_t* __restrict__ jobs, long long nb, long long nm,
                                    uint32_t* __restrict__ direct_out, int n) {
  const long long g = blockIdx.x;
  const int op = op_uniform >= 0 ? op_uniform : ops[g];
  const size_t row = (size_t)g * (n + 1);
  const long long job = job_of ? job_of[g] : g;
  uint32_t sa, sb, bias;
  if (op == 10) {
    uint32_t* o0 = jobs + (size_t)(nb + job) * (n + 1);
    uint32_t* o1 = jobs + (size_t)(nb + nm + job) * (n + 1);
    gate_coeffs(1, sa, sb, bias);
    for (int i = threadIdx.x; i <= n; i += blockDim.x)
      o0[i] = sa * a[row + i] + sb * b[row + i] + (i == n ? bias : 0u);
    gate_coeffs(6, sa, sb, bias);
    for (int i = threadIdx.x; i <= n; i += blockDim.x)
      o1[i] = sa * a[row + i] + sb * c[row + i] + (i == n ? bias : 0u);
  } else if (op >= 11) {
    gate_coeffs(op, sa, sb, bias);
    for (int i = threadIdx.x; i <= n; i += blockDim.x) direct_out[row + i] = sa * a[row + i];
  } else {
    gate_coeffs(op, sa, sb, bias);
    uint32_t* o = jobs + (size_t)job * (n + 1);
    for (int i = threadIdx.x; i <= n; i += blockDim.x)
      o[i] = sa * a[row + i] + sb * b[row + i] + (i == n ? bias : 0u);
  }
}

// Second level of MUX: out = OR(x, y) prologue = x + y + 1/8, one block per MUX gate.
__global__ void mux_or_prepare_kernel(const uint32_t* __restrict__ x, const uint32_t* __restrict__ y,
                                      uint32_t* __restrict__ out, int n) {
  const long long g = blockIdx.x;
  const size_t row = (size_t)g * (n + 1);
  for (int i = threadIdx.x; i <= n; i += blockDim.x)
    out[row + i] = x[row + i] + y[row + i] + (i == n ? 0x20000000u : 0u);
}

// Two-bootstrap MUX (opt-in, tfhe_ctx_set_mux_mode): the two AND results stay under the ring key (sample-extracted,
// no key switch: gates.bootstrapWithoutKeySwitch gates/gates.go:145-149), their sum plus 1/8 is key-switched once.
__global__ void mux_sum_kernel(const uint32_t* __restrict__ x, const uint32_t* __restrict__ y, uint32_t* __restrict__ out, int N) {
  const long long g = blockIdx.x;
  const size_t row = (size_t)g * (N + 1);
  for (int i = threadIdx.x; i <= N; i += blockDim.x) out[row + i] = x[row + i] + y[row + i] + (i == N ? 0x20000000u : 0u);
}

// gather / scatter of ciphertext rows by index list (used to compact MUX jobs)
__global__ void gather_rows_kernel(const uint32_t* __restrict__ src, const int* __restrict__ idx,
                                   uint32_t* __restrict__ dst, int words) {
  const size_t s = (size_t)idx[blockIdx.x] * words, d = (size_t)blockIdx.x * words;
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[d + i] = src[s + i];
}
__global__ void scatter_rows_kernel(const uint32_t* __restrict__ src, const int* __restrict__ idx,
                                    uint32_t* __restrict__ dst, int words) {
  const size_t d = (size_t)idx[blockIdx.x] * words, s = (size_t)blockIdx.x * words;
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[d + i] = src[s + i];
}

// ---------------------------------------------------------------------------------------------
// Levelised circuits (the caller directly above the path: README.md:78-114 full adder, gates.go:107-114 MUX).
// Wires live on the device as [wire][instance][n+1]; one launch prepares every (gate of this level) x instance.
// ---------------------------------------------------------------------------------------------
struct GateDesc { int op, in0, in1, out; };  // MUX is expanded on the host into AND / ANDNY / OR

__global__ void circuit_prepare_kernel(const GateDesc* __restrict__ gates, long long instances,
                                       const uint32_t* __restrict__ wires, uint32_t* __restrict__ prep, int n) {
  const long long j = blockIdx.x;
  const GateDesc d = gates[j / instances];
  const long long inst = j % instances;
  const uint32_t* a = wires + ((size_t)d.in0 * instances + inst) * (n + 1);
  const uint32_t* b = wires + ((size_t)d.in1 * instances + inst) * (n + 1);
  uint32_t sa, sb, bias;
  gate_coeffs(d.op, sa, sb, bias);
  uint32_t* o = prep + (size_t)j * (n + 1);
  for (int i = threadIdx.x; i <= n; i += blockDim.x) o[i] = sa * a[i] + sb * b[i] + (i == n ? bias : 0u);
}

// NOT / COPY on a whole wire (no bootstrap): out = sa * in0
__global__ void circuit_linear_kernel(GateDesc d, long long instances, uint32_t* __restrict__ wires, int n) {
  uint32_t sa, sb, bias;
  gate_coeffs(d.op, sa, sb, bias);
  const size_t total = (size_t)instances * (n + 1);
  const uint32_t* a = wires + (size_t)d.in0 * total;
  uint32_t* o = wires + (size_t)d.out * total;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) o[i] = sa * a[i];
}

// trlwe/trlwe_ops.go:10-21 with k = 0
__global__ void sample_extract_kernel(const uint32_t* __restrict__ trlwe, uint32_t* __restrict__ out, int N) {
  const long long g = blockIdx.x;
  const uint32_t* A = trlwe + (size_t)g * 2 * N;
  uint32_t* o = out + (size_t)g * (N + 1);
  for (int j = threadIdx.x; j < N; j += blockDim.x) o[j] = (j == 0) ? A[0] : ~A[N - j];
  if (threadIdx.x == 0) o[N] = A[N];
}

// ---------------------------------------------------------------------------------------------
// Identity key switch, one block per ciphertext (v1: gathers its own rows; the rows of all
// co-resident blocks come out of L2).
//   out = (0,...,0,b) - sum_{i<N, j<t, k_ij != 0} KSK[i][j][k_ij],  k_ij = digit j of a_i + 2^(31 - basebit*t)
// ksk rows are padded to `stride` words (multiple of 4) for 16-byte loads.  Row order is the
// reference's: (base*t*i + base*j + k).  Subtractions commute mod 2^32, so any order is bit-exact.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) key_switch_kernel(const uint32_t* __restrict__ lwe_in,
                                                         const uint32_t* __restrict__ ksk,
                                                         uint32_t* __restrict__ out, int N, int n, int basebit,
                                                         int t, int stride, const GateDesc* __restrict__ out_gates,
                                                         long long instances, int splits) {
  extern __shared__ uint32_t rows[];  // compacted list of non-zero row indices, capacity N*t
  __shared__ int nrows;
  const long long g = blockIdx.x;
  const uint32_t* src = lwe_in + (size_t)g * (N + 1);
  if (threadIdx.x == 0) nrows = 0;
  __syncthreads();
  const uint32_t prec = 1u << (32 - (1 + basebit * t));
  const uint32_t mask = (1u << basebit) - 1u;
  const int base = 1 << basebit;
  // Large bases (Uint sets: base = 64, key 1.57 GiB > L2): keep the rows in (i, j) order and include the k = 0 rows
  // (all-zero in the key, cloudkey.go:111; 1/base of the traffic).  Every block then walks the key in the same order,
  // so co-resident ciphertexts (each row is wanted by count/base of them) meet in L2 instead of each streaming its
  // rows from HBM.  Small bases: compact away the k = 0 rows (1/4 of them at base 4); order is irrelevant there
  // because the whole key is L2-resident.
  const bool ordered = basebit >= 4 || splits > 1;  // a split list must be the same in every block of the ciphertext
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const uint32_t abar = src[i] + prec;
    for (int j = 0; j < t; j++) {
      const uint32_t k = (abar >> (32 - (j + 1) * basebit)) & mask;
      if (ordered) rows[i * t + j] = (uint32_t)(base * t * i + base * j) + k;
      else if (k != 0) rows[atomicAdd(&nrows, 1)] = (uint32_t)(base * t * i + base * j) + k;
    }
  }
  __syncthreads();
  const int cnt_all = ordered ? N * t : nrows;
  // small batches: blockIdx.y splits the row list of one ciphertext over several blocks (a lone block would stream its
  // 20-26 MB of key rows through one SM); the slices are combined by atomic adds into a zero-initialised output row
  const int r_lo = (int)(((long long)cnt_all * blockIdx.y) / splits), cnt = (int)(((long long)cnt_all * (blockIdx.y + 1)) / splits);
  const int ncol4 = stride / 4;
  for (int c4 = threadIdx.x; c4 < ncol4; c4 += blockDim.x) {
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    int r = r_lo;
    for (; r + 8 <= cnt; r += 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++)
        v[u] = __ldg(reinterpret_cast<const uint4*>(ksk + (size_t)rows[r + u] * stride) + c4);
#pragma unroll
      for (int u = 0; u < 8; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; r < cnt; r++) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(ksk + (size_t)rows[r] * stride) + c4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    // circuits: job g = (gate of this level, instance) writes wire out_gates[gate].out of that instance
    const size_t orow = out_gates ? (size_t)out_gates[g / instances].out * instances + (size_t)(g % instances) : (size_t)g;
    uint32_t* o = out + orow * (n + 1);
    const int c = c4 * 4;
    const uint32_t bterm = (blockIdx.y == 0) ? src[N] : 0u;
    if (splits == 1) {
      if (c + 0 <= n) o[c + 0] = (c + 0 == n ? bterm : 0u) - acc.x;
      if (c + 1 <= n) o[c + 1] = (c + 1 == n ? bterm : 0u) - acc.y;
      if (c + 2 <= n) o[c + 2] = (c + 2 == n ? bterm : 0u) - acc.z;
      if (c + 3 <= n) o[c + 3] = (c + 3 == n ? bterm : 0u) - acc.w;
    } else {
      if (c + 0 <= n) atomicAdd(o + c + 0, (c + 0 == n ? bterm : 0u) - acc.x);
      if (c + 1 <= n) atomicAdd(o + c + 1, (c + 1 == n ? bterm : 0u) - acc.y);
      if (c + 2 <= n) atomicAdd(o + c + 2, (c + 2 == n ? bterm : 0u) - acc.z);
      if (c + 3 <= n) atomicAdd(o + c + 3, (c + 3 == n ? bterm : 0u) - acc.w);
    }
  }
}

// zero the output rows of a split key switch (job -> wire mapping as in key_switch_kernel)
__global__ void zero_out_rows_kernel(uint32_t* __restrict__ out, int n, const GateDesc* __restrict__ out_gates, long long instances) {
  const long long g = blockIdx.x;
  const size_t orow = out_gates ? (size_t)out_gates[g / instances].out * instances + (size_t)(g % instances) : (size_t)g;
  for (int w = threadIdx.x; w <= n; w += blockDim.x) out[orow * (n + 1) + w] = 0u;
}

// ksk [rows][n+1] -> [rows][stride], zero padded.  Rows with digit k = 0 (row index a multiple of base) are never read by
// the reference (trgsw/keyswitch.go:30 `if k != 0`); they are stored as zeros so that every evaluation order of the
// key switch (compacted gather, ordered split gather, tensor-core contraction) agrees whatever the caller uploaded there.
__global__ void ksk_repack_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int n1, int stride, int base) {
  const size_t r = blockIdx.x;
  const bool k0 = (r % (size_t)base) == 0;
  for (int i = threadIdx.x; i < stride; i += blockDim.x) dst[r * stride + i] = (i < n1 && !k0) ? src[r * n1 + i] : 0u;
}

// Bootstrapping key: reference FourierPoly layout -> engine layout.
//   src: [polys][N] doubles, poly = (i*2L + r)*2 + ab, groups of (4 re, 4 im): complex index k has
//        re at (k/4)*8 + k%4 and im at (k/4)*8 + 4 + k%4   (poly/poly.go:54-62)
//   dst: [polys][8][T] double2, spectrum position k = 8*tau + e stored at [e][tau], scaled by 1/M
//        (the inverse transform's 1/(N/2), fourier_transform.go:318-346; a power of two, exact).
__global__ void bsk_repack_kernel(const double* __restrict__ src, double2* __restrict__ dst, int N, int L, int bgbit) {
  const int M = N / 2, T = M / 8;
  const size_t poly = blockIdx.x;  // ((step * 2L + row) * 2 + {A,B})
  // 2/N: the inverse transform's scale.  2^-sh: the digits of level (row mod L) reach the transform scaled by 2^sh
  // (digit_scaled, blind_rotate.cuh); both are powers of two, so every product keeps the reference's exact value.
  const int lvl = (int)((poly / 2) % (size_t)(2 * L)) % L;
  double scale = 1.0 / (double)M;
  for (int k = 0; k < 32 - (lvl + 1) * bgbit; k++) scale *= 0.5;
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    const double re = src[poly * N + (k >> 2) * 8 + (k & 3)];
    const double im = src[poly * N + (k >> 2) * 8 + 4 + (k & 3)];
    const int tau = k >> 3, e = k & 7;
    const size_t rowset = poly / (size_t)(4 * L);                 // CMUX step
    const int r = (int)((poly / 2) % (size_t)(2 * L)), ab = (int)(poly & 1);
#if TFHE_BR_KEY256
    const size_t pos = ((size_t)(r * 8 + e) * T + tau) * 2 + ab;  // key_pos (blind_rotate.cuh): A and B values adjacent
#else
    const size_t pos = ((size_t)(r * 2 + ab) * 8 + e) * T + tau;
#endif
    dst[rowset * (size_t)(4 * L) * M + pos] = make_double2(re * scale, im * scale);
  }
}

}  // namespace tfhe
