"""Round-2 hot-path invocations for compute-sanitizer (memcheck / racecheck / synccheck), kept short by forcing most CMUX
steps onto the skip path: the persistent work-item kernel with a forced hand-over every 8 steps, the latency kernels, the
LUT-index / many-LUT / two-rotation-MUX options, the pipelined host path with tiny chunks, re-encryption, device key
generation and all three key-switch evaluations."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")


def short(ct, P, live=24):
    ct = ct.copy()
    ct[:, live:P.n] = 0      # a~ = 0 => skipped steps
    return ct


P = T.params.get("80")
sk = T.key.NewSecretKey(P, 1)
ctx = T.Context(P, 0)
ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=5, with_ksk=True, export=False)   # device key generation kernels
ct = short(T.tlwe.EncryptBool([0, 1, 1, 0, 1], sk, 3), P)
ref = ctx.gate_batch("NAND", ct, ct)                               # latency kernel (5 gates)
ctx.set_blind_rotate_variant(10)                                   # throughput kernel at every batch size
ctx.set_blind_rotate_chunk_steps(8)                                # 69 work items per gate, accumulator handed over each time
assert np.array_equal(ctx.gate_batch("NAND", ct, ct), ref)
assert ctx.blind_rotate_batch(ct).size == 5 * 2 * P.N
ctx.set_pipeline_chunk(2)                                          # 3 chunks through the two staging slots
assert np.array_equal(ctx.gate_batch(["NAND"] * 5, ct, ct), ref)
ctx.gate_batch(["MUX", "NOT", "XOR", "COPY", "AND"], ct, ct, ct)
ctx.set_mux_mode(1)
ctx.gate_batch(["MUX", "MUX", "XOR", "MUX", "AND"], ct, ct, ct)
ctx.set_mux_mode(0)
ctx.set_pipeline_chunk(16384)
m = 2
msgs = np.arange(5) % m
cm = short(T.tlwe.EncryptLWEMessage(msgs, m, sk, 7), P)
gen = T.lut.NewGenerator(m, P)
luts = np.stack([gen.GenLookUpTable(f).Poly.reshape(-1) for f in (lambda v: v, lambda v: 1 - v)])
ctx.bootstrap_batch_indexed(cm, luts, np.array([0, 1, 1, 0, 1]))
ctx.bootstrap_multi_lut_batch(cm, T.lut.PackLookUpTables(list(luts)).reshape(1, -1), 1)
ctx.set_blind_rotate_variant(0)
ctx.set_blind_rotate_chunk_steps(0)
ext = np.random.default_rng(0).integers(0, 1 << 32, (200, P.N + 1), dtype=np.uint64).astype(np.uint32)
for v in ("gather", "mma"):
    ctx.set_key_switch_variant(v)
    ctx.key_switch_batch(ext)
ctx.set_key_switch_variant("auto")
key = np.random.default_rng(1).integers(0, 1 << 32, (P.n * 2 * 4, P.n + 1), dtype=np.uint64).astype(np.uint32)
ctx.load_reencryption_key(key, 2, 2)
ctx.reencrypt_batch(ct)
ctx.close()

P2 = T.params.get("uint2")                                         # L = 1 kernels, order-preserving latency kernel, tiled key switch
sk2 = T.key.NewSecretKey(P2, 11)
c2 = T.Context(P2, 0)
c2.generate_cloudkey(sk2.KeyLv0, sk2.KeyLv1, seed=6, with_ksk=True, export=False)
cm2 = short(T.tlwe.EncryptLWEMessage(np.arange(3) % 4, 4, sk2, 8), P2)
lut2 = T.lut.NewGenerator(4, P2).GenLookUpTable(lambda v: 3 - v).Poly.reshape(1, -1)
c2.bootstrap_batch(cm2, lut2)
ext2 = np.random.default_rng(2).integers(0, 1 << 32, (300, P2.N + 1), dtype=np.uint64).astype(np.uint32)
for v in ("gather", "tile"):
    c2.set_key_switch_variant(v)
    c2.key_switch_batch(ext2)
c2.close()
print("sanitize workload done")
