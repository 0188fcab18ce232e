"""A few programmable bootstraps of one parameter set at one batch size (ncu target): python tools/pbs_run.py <set> <count> [reps] [auto|gather|tile]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")
name, count = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
P = T.params.get(name)
m = {"uint1": 2, "uint2": 4, "uint3": 8, "uint4": 16, "uint5": 32}.get(name, 2)
sk = T.key.NewSecretKey(P, 1)
ctx = T.Context(P, 0)
ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=2, with_ksk=True, export=False)
msgs = np.arange(count) % m
ct = T.tlwe.EncryptLWEMessage(msgs, m, sk, 3)
lut = T.lut.NewGenerator(m, P).GenLookUpTable(lambda v: (m - 1) - v).Poly.reshape(1, -1)
if len(sys.argv) > 4:
    ctx.set_key_switch_variant(sys.argv[4])
ctx.set_timing(True)
for _ in range(reps):
    out = ctx.bootstrap_batch(ct, lut)
    tm = ctx.collect_timing()
print(name, count, "correct", bool(np.array_equal(T.tlwe.DecryptLWEMessage(out, m, sk), (m - 1) - msgs)), tm)
