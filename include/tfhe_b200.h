/*
 * tfhe_b200.h — C ABI of the B200-native TFHE gate-bootstrap engine.
 *
 * This is the drop-in boundary for the hot path of thedonutfactory/go-tfhe (reference paths
 * below are relative to the reference repository).  The reference has no FFI of its own — it
 * is a single Go module — so the boundary is cut at the narrowest waist of its call graph:
 * flattened ciphertext batches in, flattened ciphertext batches out, cloud key resident on the
 * device.  A thin cgo shim (go/) flattens the reference's pointer-rich Go
 * types into these buffers and keeps the gates.* / evaluator.* signatures unchanged.
 *
 * Conventions
 *   - All functions return 0 on success and a negative tfhe_status on failure; the message is
 *     available from tfhe_last_error().  (The reference panics; the Go shim re-panics.)
 *   - Host buffers are caller-owned, contiguous, and borrowed only for the duration of a call.
 *   - One call at a time per context.  One context per GPU (one process per GPU in multi-GPU
 *     runs; gates shard by index, keys are replicated, no collective on the hot path).
 *   - There is no CPU fallback: every entry point fails with TFHE_ERR_CUDA if no sm_100 device.
 *
 * Flattened layouts (Torus = uint32_t, reference params/params.go:27)
 *   LWE ciphertext batch   [count][n+1]          tlwe.TLWELv0.P               tlwe/tlwe.go:11-13
 *   TRLWE / LUT batch      [count][2][N]         trlwe.TRLWELv1 {A,B}         trlwe/trlwe.go:13-16
 *   key-switching key      [N][t][base][n+1]     CloudKey.KeySwitchingKey, row index
 *                                                base*t*i + base*j + k        cloudkey/cloudkey.go:111
 *   bootstrapping key      [n][2L][2][N] double  CloudKey.BootstrappingKey[i].TRLWEFFT[r].{A,B}
 *                                                in the reference FourierPoly layout (groups of
 *                                                4 real then 4 imaginary parts)
 *                                                trgsw/trgsw.go:60-68, poly/poly.go:54-62
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tfhe_ctx tfhe_ctx;

/* Mirrors params.TLWELv0Params.N and params.TRGSWLv1Params{N,L,BGBIT,BASEBIT,IKS_T}
 * (params/params.go:50-78).  Supported shapes: N in {512,1024,2048}, L*bgbit <= 32. */
typedef struct {
  int32_t n;       /* LWE dimension (TLWELv0.N)            */
  int32_t N;       /* ring degree   (TRGSWLv1.N)           */
  int32_t L;       /* gadget levels (TRGSWLv1.L)           */
  int32_t bgbit;   /* log2 gadget base (TRGSWLv1.BGBIT)    */
  int32_t basebit; /* log2 key-switch base (BASEBIT)       */
  int32_t iks_t;   /* key-switch levels (IKS_T)            */
} tfhe_params;

typedef enum {
  TFHE_OK = 0,
  TFHE_ERR_ARG = -1,     /* bad argument / unsupported parameter shape */
  TFHE_ERR_CUDA = -2,    /* CUDA runtime failure or no usable device   */
  TFHE_ERR_STATE = -3,   /* e.g. cloud key not loaded                  */
  TFHE_ERR_NOMEM = -4
} tfhe_status;

/* Gate opcodes for tfhe_gate_batch.  Linear prologue c = sa*a + sb*b + (0,..,0,bias), then one
 * bootstrap with the default test vector.  evaluator/gates_helper.go:10-63 (NAND, AND, OR, XOR),
 * gates/gates.go:52-104 (XNOR, NOR, ANDNY, ANDYN, ORNY, ORYN), gates/gates.go:107-114 (MUX =
 * OR(AND(a,b), AND(NOT a, c)), three bootstraps), gates/gates.go:117-130 (NOT, COPY: no bootstrap). */
typedef enum {
  TFHE_OP_NAND = 0, TFHE_OP_AND = 1, TFHE_OP_OR = 2, TFHE_OP_XOR = 3, TFHE_OP_XNOR = 4,
  TFHE_OP_NOR = 5, TFHE_OP_ANDNY = 6, TFHE_OP_ANDYN = 7, TFHE_OP_ORNY = 8, TFHE_OP_ORYN = 9,
  TFHE_OP_MUX = 10, TFHE_OP_NOT = 11, TFHE_OP_COPY = 12
} tfhe_op;

/* --- lifecycle ------------------------------------------------------------------------------ */

/* Creates a context on CUDA device `device` (>= 0).  Replaces evaluator.NewEvaluator
 * (evaluator/evaluator.go:27-35) and the package-global evaluator of gates/gates.go:19-23. */
int tfhe_ctx_create(const tfhe_params* params, int device, tfhe_ctx** out);
/* One context over `ndev` GPUs of this process (devices[0..ndev-1]; devices == NULL: 0..ndev-1; ndev <= 0: every
 * visible GPU).  The cloud key is uploaded or generated once and replicated to the other devices by peer copies
 * (NVLink / NVSwitch); the host-buffer batch calls (tfhe_gate_batch, tfhe_bootstrap_batch, tfhe_blind_rotate_batch,
 * tfhe_circuit_run) shard their batch by index (circuits: by instance) over the devices, one host thread per device,
 * nothing crosses devices on the hot path.  This is what the goroutine fan-out of trgsw.BatchBlindRotate
 * (trgsw/trgsw.go:234-252) and gates.Batch* (gates/gates.go:156-312) becomes on a multi-GPU node; results are
 * identical to a single-device context.  The *_device entry points need a single-device context. */
int tfhe_ctx_create_multi(const tfhe_params* params, int ndev, const int* devices, tfhe_ctx** out);
/* Number of GPUs behind `ctx` (1 for tfhe_ctx_create). */
int tfhe_ctx_device_count(const tfhe_ctx* ctx);
void tfhe_ctx_destroy(tfhe_ctx* ctx);
/* Message of the last failure on `ctx` (or of the last failed tfhe_ctx_create if ctx == NULL). */
const char* tfhe_last_error(const tfhe_ctx* ctx);

/* Uploads a cloudkey.CloudKey (cloudkey/cloudkey.go:16-21).  bsk_fft is taken in the reference's
 * own Fourier layout and repacked on the device (scaled by 2/N, an exact power of two) into the
 * engine's layout; ksk rows are re-strided for 16-byte loads; ksk may be NULL if only
 * tfhe_blind_rotate_batch is used (cloudkey.NewCloudKeyNoKSK, cloudkey.go:34-57). */
int tfhe_ctx_load_cloudkey(tfhe_ctx* ctx, uint32_t decomposition_offset, const double* bsk_fft,
                           const uint32_t* ksk, const uint32_t* testvec);
/* cloudkey.NewCloudKey (cloudkey/cloudkey.go:24-145) evaluated on the device from the caller's secret key
 * (key.SecretKey.KeyLv0 [n], KeyLv1 [N], binary, as u32): genBootstrappingKey (:122-145: trgsw.EncryptTorus
 * trgsw/trgsw.go:32-58 over trlwe.EncryptF64 trlwe/trlwe.go:28-50, then NewTRGSWLv1FFT :72-82), genKeySwitchingKey
 * (:88-120, tlwe.EncryptF64 tlwe/tlwe.go:36-52 with alpha_lv0 = params.KSKAlpha()), genTestvec, genDecompositionOffset.
 * alpha_lv1 = params.BSKAlpha().  Randomness is ChaCha20 under a 256-bit key, public masks and secret noise in separate
 * streams: seed == 0 takes that key from the operating system's entropy source (getrandom) — the setting for real keys,
 * like the reference's self-seeding math/rand but cryptographically strong; seed != 0 expands the seed into the key so
 * that (secret key, seed) -> cloud key is reproducible (tests, benchmarks; 64 bits of entropy only).  The key is left
 * loaded in ctx; the optional outputs receive the CloudKey fields in the layouts of tfhe_ctx_load_cloudkey (any of them
 * may be NULL; ksk_out requires with_ksk != 0).  The secret key is only read by this call and never stored. */
int tfhe_ctx_generate_cloudkey(tfhe_ctx* ctx, const uint32_t* key_lv0, const uint32_t* key_lv1, double alpha_lv0,
                               double alpha_lv1, uint64_t seed, int with_ksk, uint32_t* decomposition_offset_out,
                               double* bsk_fft_out, uint32_t* ksk_out, uint32_t* testvec_out);
/* Same, from buffers already in this device's memory (e.g. filled by one NCCL broadcast from
 * rank 0).  `stream` is a cudaStream_t (0 = default stream); returns after the repack is done. */
int tfhe_ctx_load_cloudkey_device(tfhe_ctx* ctx, uint32_t decomposition_offset, const double* d_bsk_fft,
                                  const uint32_t* d_ksk, const uint32_t* d_testvec, void* stream);

/* --- the hot path, host buffers --------------------------------------------------------------- */

/* count independent bootstraps: blind rotate -> sample extract(0) -> identity key switch.
 * Replaces Evaluator.BootstrapAssign (evaluator/evaluator.go:139-148) and, with luts != NULL,
 * Evaluator.BootstrapLUTAssign (evaluator/programmable_bootstrap.go:93-115), applied to every
 * element.  luts: NULL => CloudKey.BlindRotateTestvec for all; else nluts in {1, count} TRLWE
 * test vectors [nluts][2][N] (lut.LookUpTable.Poly). */
int tfhe_bootstrap_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* ct_in, const uint32_t* luts,
                         int64_t nluts, uint32_t* ct_out);

/* The same with a LUT TABLE and one index per ciphertext: luts [nluts][2][N], lut_index [count] with values in
 * [0, nluts) (host memory).  What BootstrapLUTAssign callers that reuse a handful of functions want
 * (examples/add_two_numbers/main.go:59-72 builds three LUTs for every nibble): 16 KiB per LUT cross the bus instead of
 * 16 KiB per ciphertext.  Results equal tfhe_bootstrap_batch with the LUTs expanded. */
int tfhe_bootstrap_batch_indexed(tfhe_ctx* ctx, int64_t count, const uint32_t* ct_in, const uint32_t* luts,
                                 int64_t nluts, const int32_t* lut_index, uint32_t* ct_out);

/* Many-LUT programmable bootstrap (additive; no reference counterpart): 2^log2_k functions of every ciphertext from ONE
 * blind rotation.  The mod switch goes to 2N / 2^log2_k levels (scaled back, so every rotation is a multiple of 2^log2_k),
 * the test vector interleaves the functions (packed[j] = lut_{j mod k}[j], k = 2^log2_k, all built by
 * lut.Generator.GenLookUpTable for the same message modulus), and after the rotation the samples at indices 0..k-1
 * (trlwe.SampleExtractIndex, trlwe/trlwe.go:114-128) are key-switched: ct_out [count][k][n+1], output i decrypts to
 * f_i(message).  Costs log2_k bits of the mod-switch precision (the usual many-LUT trade).  packed_luts: nluts in
 * {1, count} test vectors [nluts][2][N]. */
int tfhe_bootstrap_multi_lut_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* ct_in, const uint32_t* packed_luts,
                                   int64_t nluts, int32_t log2_k, uint32_t* ct_out);

/* count gates.  Replaces gates.{NAND..ORYN,MUX,NOT,Copy} (gates/gates.go:26-130) and
 * gates.Batch{NAND,AND,OR,XOR,NOR,XNOR} (gates/gates.go:156-312; every element equals the
 * single-gate path, and XNOR uses the single-gate bias, see SURVEY.md section 2 defects 1-2).
 * ops: nops in {1, count} opcodes.  c is read only by TFHE_OP_MUX gates and may be NULL otherwise;
 * b is ignored by NOT / COPY. */
int tfhe_gate_batch(tfhe_ctx* ctx, int64_t count, const uint8_t* ops, int64_t nops, const uint32_t* a,
                    const uint32_t* b, const uint32_t* c, uint32_t* out);

/* Blind rotation only; output TRLWE [count][2][N].  Replaces Evaluator.BlindRotateAssign
 * (evaluator/evaluator.go:110-135), trgsw.BlindRotate / trgsw.BatchBlindRotate
 * (trgsw/trgsw.go:197-252). */
int tfhe_blind_rotate_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* ct_in, const uint32_t* luts,
                            int64_t nluts, uint32_t* trlwe_out);

/* CMUX with bootstrapping-key row `bsk_index`: out = ct0 + BK[bsk_index] (x) (ct1 - ct0).
 * Replaces Evaluator.CMuxAssign / ExternalProductAssign (evaluator/evaluator.go:50-106),
 * trgsw.CMUX / ExternalProductWithFFT (trgsw/trgsw.go:108-194).  ct0 == NULL means ct0 = 0, i.e.
 * the plain external product of ct1.  TRLWE batches [count][2][N]. */
int tfhe_cmux_batch(tfhe_ctx* ctx, int64_t count, int32_t bsk_index, const uint32_t* ct0,
                    const uint32_t* ct1, uint32_t* out);

/* Sample extract at index 0 (trlwe/trlwe_ops.go:10-21): TRLWE [count][2][N] -> LWE [count][N+1]. */
int tfhe_sample_extract_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* trlwe_in, uint32_t* lwe_out);

/* Identity key switch N -> n (trgsw/keyswitch.go:10-37, trgsw/trgsw.go:285-311):
 * LWE [count][N+1] -> LWE [count][n+1]. */
int tfhe_key_switch_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* lwe_in, uint32_t* ct_out);

/* --- proxy re-encryption (proxyreenc/proxyreenc.go) on the key-switch kernel ------------------------------------------- */
/* Uploads proxyreenc.ProxyReencryptionKey.KeyEncryptions flattened as [n*t*base][n+1] (row index base*t*i + base*j + k,
 * proxyreenc.go:272-290; k = 0 rows are never read), with its Base = 2^basebit and T = t. */
int tfhe_ctx_load_reencryption_key(tfhe_ctx* ctx, const uint32_t* key_encryptions, int32_t basebit, int32_t t);
/* proxyreenc.ReencryptTLWELv0 (proxyreenc.go:321-366) for every ciphertext: [count][n+1] under the source key ->
 * [count][n+1] under the target key.  Bit-identical to the reference (u32 arithmetic). */
int tfhe_reencrypt_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* ct_in, uint32_t* ct_out);

/* --- polynomial transforms at API granularity (reference FourierPoly layout in and out) ---------------------- */
/* poly.Evaluator.ToFourierPolyAssign (poly/fourier_transform.go:18-21): [count][N] u32 -> [count][N] f64, unscaled,
 * coefficients read as int32, output in the reference's order and 4-real/4-imaginary packing. */
int tfhe_to_fourier_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* poly_in, double* fourier_out);
/* poly.Evaluator.ToPolyAssign (poly/fourier_transform.go:31-44): inverse transform, divide by N/2, reduce mod 2^32
 * with rounding, [count][N] f64 -> [count][N] u32. */
int tfhe_to_poly_batch(tfhe_ctx* ctx, int64_t count, const double* fourier_in, uint32_t* poly_out);
/* poly.Evaluator.MulPolyAssign (poly/poly_mul.go:12-22): negacyclic product of two torus polynomials, both read as
 * int32 coefficients, [count][N] x [count][N] -> [count][N]. */
int tfhe_mul_poly_batch(tfhe_ctx* ctx, int64_t count, const uint32_t* p0, const uint32_t* p1, uint32_t* out);

/* --- levelised circuits (additive: the caller immediately above the path) ------------------------------- */
/* One gate of a circuit over wire ids.  Wires 0..n_inputs-1 are the inputs; every gate writes a distinct wire
 * >= n_inputs and may read only inputs or wires written by EARLIER gates of the list (topological order).
 * in2 is read by TFHE_OP_MUX only; in1 is ignored by NOT / COPY. */
typedef struct { uint8_t op; int32_t in0, in1, in2, out; } tfhe_gate_desc;

/* Evaluates the same circuit on `instances` independent input sets.  The engine levelises the gate list
 * (depth = 1 + max depth of the operands; NOT/COPY cost no level; MUX = OR(AND(a,b), AND(NOT a, c)) as in
 * gates/gates.go:107-114) and runs each level as ONE batch of (gates of the level) x instances bootstraps with all
 * intermediate wires resident on the device.  Every gate computes exactly what gates.X of the reference computes
 * (e.g. the full adder of README.md:78-87), so results are bit-identical to running the gates one by one.
 * inputs: [n_inputs][instances][n+1]; outputs: [n_outputs][instances][n+1] (wire output_wires[k]). */
int tfhe_circuit_run(tfhe_ctx* ctx, int64_t instances, int32_t n_inputs, int32_t n_gates, const tfhe_gate_desc* gates,
                     const uint32_t* inputs, int32_t n_outputs, const int32_t* output_wires, uint32_t* outputs);

/* --- the hot path, device buffers (inputs already resident in HBM) -------------------------- */
/* Same semantics; ciphertext / LUT pointers are device pointers on the context's device; work is enqueued on
 * `stream` (cudaStream_t, 0 = default) and NOT synchronised: the call returns as soon as everything is enqueued.
 * `ops` of tfhe_gate_batch_device is a HOST array (nops in {1, count} opcodes), read before the call returns.
 * d_out may be the same buffer as d_a, d_b or d_c (prepared ciphertexts go to internal scratch, never to d_out). */
int tfhe_bootstrap_batch_device(tfhe_ctx* ctx, int64_t count, const uint32_t* d_ct_in, const uint32_t* d_luts,
                                int64_t nluts, uint32_t* d_ct_out, void* stream);
int tfhe_gate_batch_device(tfhe_ctx* ctx, int64_t count, const uint8_t* ops, int64_t nops, const uint32_t* d_a,
                           const uint32_t* d_b, const uint32_t* d_c, uint32_t* d_out, void* stream);

/* --- introspection ---------------------------------------------------------------------------- */
/* Number of CUDA kernels this context has launched since creation (bench.py: gpu_launches). */
int64_t tfhe_ctx_kernel_launches(const tfhe_ctx* ctx);
/* Selects the blind-rotate kernel.  0 = automatic (default): the persistent throughput kernel (one 64-thread block per
 * gate-item, key rows by LDG straight from L2), except for small batches — at most two gates per SM on the 80/110/128-bit
 * sets run kernel 9, at most one ciphertext per SM on the L <= 2 sets (Uint1-5 / programmable bootstraps) runs kernel 12.
 * 9 = latency mode: four warps per gate, the two polynomials of a CMUX step in parallel (exact sets only).
 * 10 = the throughput kernel at every batch size.  12 = latency mode that keeps the reference's accumulation order
 * (bit-identical for every set).  All compute identical results.  Values 1-8, 11, 13 name the measured-slower round-1
 * experiments (profiles/r01_experiments.md) and exist only in a library built with -DTFHE_EXPERIMENTAL=1. */
int tfhe_ctx_set_blind_rotate_variant(tfhe_ctx* ctx, int variant);
/* How TFHE_OP_MUX is evaluated.  0 (default) = exactly gates.MUX (gates/gates.go:107-114): OR(AND(a,b), AND(NOT a, c)),
 * three bootstraps, bit-identical to the reference.  1 (opt-in) = the two ANDs are blind-rotated and sample-extracted
 * without key switch (gates.bootstrapWithoutKeySwitch, gates.go:145-149), summed with 1/8 and key-switched once: two blind
 * rotations instead of three; same truth table, but NOT the reference's ciphertext words (hence opt-in). */
int tfhe_ctx_set_mux_mode(tfhe_ctx* ctx, int mode);
/* tfhe_circuit_run, CUDA-graph replay (off by default).  When enabled, the first call with a given (gate list, n_inputs,
 * instances) runs as usual — which also sizes every scratch buffer —, the second call records the same per-level
 * launches into a CUDA graph, and later calls replay that graph with one launch, for as long as no device buffer of the
 * library has been reallocated since (otherwise it is recorded again).  Results are identical; what it removes is the
 * per-launch host cost (~100 launches for an 8-bit adder).  tfhe_ctx_circuit_graph_replays counts the replays. */
int tfhe_ctx_set_circuit_graph(tfhe_ctx* ctx, int enable);
int64_t tfhe_ctx_circuit_graph_replays(const tfhe_ctx* ctx);
/* The throughput kernel runs persistent blocks over WORK ITEMS of `steps` consecutive CMUX steps of one gate (the
 * accumulator is handed from item to item through device memory), so that a batch that is not a multiple of the
 * resident blocks still fills the SMs to the end.  0 = automatic (default: whole gates for batches that fit the
 * resident blocks, ~n/14-step items above), >= n = whole gates.  Results do not depend on it. */
int tfhe_ctx_set_blind_rotate_chunk_steps(tfhe_ctx* ctx, int steps);
/* Host-buffer batch calls larger than 1.5 chunks of `rows` ciphertexts (default 16384) are pipelined: the transfers of
 * chunk k+1 (in) and k-1 (out) overlap the kernels of chunk k on separate streams, and device staging stays bounded by
 * two chunks whatever the batch size.  Results do not depend on it. */
int tfhe_ctx_set_pipeline_chunk(tfhe_ctx* ctx, int64_t rows);
/* Selects how IdentityKeySwitching (trgsw/keyswitch.go:10-37) is evaluated: 0 = automatic (default: the contraction wherever it exists, tiles for large bases and batches), 1 = one block per
 * ciphertext gathering its N*t*(1-1/base) key rows out of L2, 2 = the whole batch as one exact u8 x u8 -> s32
 * contraction on the tensor cores (tcgen05.mma kind::i8 over the byte planes of the key; basebit = 2 parameter sets
 * only), 3 = shared-memory tiles for the basebit >= 4 sets (Uint2-5): a block owns 256 ciphertexts x 64 output words and
 * stages the `base` candidate rows of each (i, j) once for all of them (automatic from 160 ciphertexts on).  All are
 * bit-identical: additions mod 2^32 commute. */
int tfhe_ctx_set_key_switch_variant(tfhe_ctx* ctx, int variant);
/* Per-stage device timing for bench.py's roofline: when enabled, every bootstrap batch records CUDA events on
 * its launching stream around the blind-rotate kernel and the key-switch kernel.  tfhe_ctx_collect_timing waits
 * for the recorded events and returns {blind_rotate_ms_total, blind_rotate_launches, key_switch_ms_total,
 * key_switch_launches} since the previous collect. */
int tfhe_ctx_set_timing(tfhe_ctx* ctx, int enable);
int tfhe_ctx_collect_timing(tfhe_ctx* ctx, double out[4]);
/* Algorithmic bytes per bootstrap of SURVEY.md section 8(d): n*2L*2*N*8 + N*t*(1-1/base)*(n+1)*4 + io. */
int64_t tfhe_ctx_algorithmic_bytes_per_bootstrap(const tfhe_ctx* ctx);
/* Measured FP64 roof of `device` (bench.py roofline_fp64): independent DFMA chains with reuse-cache operands, the
 * friendliest instruction mix.  out = {TFLOP/s at 2 flops per DFMA, thread-DFMA per clock per SM at the driver-reported
 * SM clock, that clock in MHz}.  Takes ~0.1 s. */
int tfhe_fp64_peak_probe(int device, double out[3]);
/* Library version string. */
const char* tfhe_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TFHE_B200_H */
