"""Small hot-path invocations for compute-sanitizer (memcheck / racecheck): one CMUX step batch, a short blind
rotation (first 24 mask words non-zero, the rest forced to the skip path) and a key switch, for every kernel variant."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")
P = T.params.get("80")
sk = T.key.NewSecretKey(P, 1)
ck = T.cloudkey.NewCloudKey(sk, 2)
ctx = ck.engine(0)
rng = np.random.default_rng(0)
c0 = rng.integers(0, 1 << 32, (3, 2 * P.N), dtype=np.uint64).astype(np.uint32)
c1 = rng.integers(0, 1 << 32, (3, 2 * P.N), dtype=np.uint64).astype(np.uint32)
ctx.cmux_batch(0, c0, c1)
ct = T.tlwe.EncryptBool([0, 1, 1], sk, 3)
ct[:, 24:P.n] = 0  # a~ = 0 => skipped steps: keeps the sanitizer run short
ref = None
only = [a.split("=")[1] for a in sys.argv if a.startswith("--variant=")]
for v in (("ldg",) if "--default-only" in sys.argv else (only or ("ldg", "tma", "tex", "w16", "tmem", "tmex", "tmex+tma"))):
    ctx.set_blind_rotate_variant(v)
    out = ctx.blind_rotate_batch(ct)
    ref = out if ref is None else ref
    assert np.array_equal(out, ref), v
ctx.set_blind_rotate_variant("ldg")
ext = ctx.sample_extract_batch(ref)
ks = ctx.key_switch_batch(ext)
ctx.set_key_switch_variant("mma")   # tensor-core key switch (TMA + tcgen05 + TMEM), partial tiles
ext200 = np.tile(ext, (67, 1))[:200]
assert np.array_equal(ctx.key_switch_batch(ext200)[:3], ks)
ctx.set_key_switch_variant("auto")
ctx.gate_batch(["MUX", "NOT", "XOR"], ct, ct, ct)
ck2 = T.cloudkey.NewCloudKeyOnDevice(sk, 5)   # device key generation kernels
assert list(T.tlwe.DecryptBool(T.gates.NAND(ct[:1] * 0 + T.tlwe.EncryptBool([1], sk, 9), T.tlwe.EncryptBool([1], sk, 10), ck2), sk)) == [0]
ck2.close()
print("sanitize workload done")
