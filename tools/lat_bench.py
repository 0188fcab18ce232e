"""Call latency of small batches through the C ABI (host buffers): python tools/lat_bench.py"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")
for name, counts in (("80", (1, 148, 296)), ("128", (1, 148, 296)), ("uint5", (1, 148)), ("uint3", (1,))):
    P = T.params.get(name)
    sk = T.key.NewSecretKey(P, 1)
    ctx = T.Context(P, 0)
    ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=2, with_ksk=True, export=False)
    ctx.set_timing(True)
    for c in counts:
        if name.startswith("uint"):
            m = {"uint3": 8, "uint5": 32}[name]
            ct = T.tlwe.EncryptLWEMessage(np.arange(c) % m, m, sk, 3)
            lut = T.lut.NewGenerator(m, P).GenLookUpTable(lambda v: v).Poly.reshape(1, -1)
            call = lambda: ctx.bootstrap_batch(ct, lut)
        else:
            a = T.tlwe.EncryptBool(np.arange(c) % 2, sk, 3)
            call = lambda: ctx.gate_batch("NAND", a, a)
        for _ in range(3):
            call()
        ctx.collect_timing()
        t0 = time.perf_counter()
        for _ in range(10):
            call()
        dt = (time.perf_counter() - t0) / 10
        tm = ctx.collect_timing()
        print("%-6s count %4d  call %.3f ms  BR kernel %.3f ms" % (name, c, dt * 1e3, tm["blind_rotate_ms"] / tm["blind_rotate_launches"]))
    ctx.close()
