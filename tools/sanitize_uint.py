"""compute-sanitizer workload for the L <= 2 path: order-preserving latency kernel + split key-switch gather (Uint2)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")
P = T.params.get("uint2")
sk = T.key.NewSecretKey(P, 1)
ck = T.cloudkey.NewCloudKeyOnDevice(sk, 2, export=False)
ctx = ck.engine(0)
ct = T.tlwe.EncryptLWEMessage(np.array([0, 1, 2, 3]), 4, sk, 3)
ct[:, 20:P.n] = 0  # skipped steps keep the run short
lut = T.lut.NewGenerator(4, P).GenLookUpTable(lambda x: (x + 1) % 4).Poly.reshape(1, -1)
ref = None
for v in ("throughput", "latp"):
    ctx.set_blind_rotate_variant(v)
    out = ctx.bootstrap_batch(ct, lut)
    ref = out if ref is None else ref
    assert np.array_equal(out, ref), v
print("sanitize (uint2) workload done")
