"""Soak of the persistent work-item kernel: many launches at random batch sizes / parameter sets / item lengths, every
result checked by decryption.  python tools/soak.py [seconds]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(123)
ctxs = []
for name in ("80", "128", "uint3"):
    P = T.params.get(name)
    sk = T.key.NewSecretKey(P, 5)
    ctx = T.Context(P, 0)
    ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=6, with_ksk=True, export=False)
    bits = rng.integers(0, 2, 6000).astype(np.uint8)
    ctxs.append((name, P, sk, ctx, bits, T.tlwe.EncryptBool(bits, sk, 7), T.tlwe.EncryptBool(1 - bits, sk, 8)))
t0, launches, gates = time.time(), 0, 0
while time.time() - t0 < budget:
    name, P, sk, ctx, bits, a, b = ctxs[rng.integers(0, len(ctxs))]
    count = int(rng.choice([297, 593, 700, 1024, 1500, 2048, 3000, 4096, 5000, int(rng.integers(300, 6000))]))
    ctx.set_blind_rotate_chunk_steps(int(rng.choice([0, 0, 0, 17, 54, 200])))
    op = ["NAND", "XOR", "AND"][int(rng.integers(0, 3))]
    out = ctx.gate_batch(op, a[:count], b[:count])
    x, y = bits[:count], 1 - bits[:count]
    want = {"NAND": 1 - (x & y), "XOR": x ^ y, "AND": x & y}[op]
    assert np.array_equal(T.tlwe.DecryptBool(out, sk), want), (name, count, op)
    launches += 1
    gates += count
print("soak ok: %d calls, %d gates, %.1f s" % (launches, gates, time.time() - t0))
