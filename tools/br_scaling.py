"""Blind-rotation kernel time against batch size (128-bit NAND gates): python tools/br_scaling.py [counts...]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")
P = T.params.get("128")
sk = T.key.NewSecretKey(P, 1)
ctx = T.Context(P, 0)
ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=2, with_ksk=True, export=False)
counts = [int(a) for a in sys.argv[1:]] or [592, 1184, 2048, 2368, 4096, 4736, 8192, 16384]
bits = np.arange(max(counts)) % 2
ct = T.tlwe.EncryptBool(bits, sk, 3)
ctx.set_blind_rotate_variant(10)
ctx.set_timing(True)
for c in counts:
    for _ in range(2):
        ctx.gate_batch("NAND", ct[:c], ct[:c])
    ctx.collect_timing()
    for _ in range(3):
        ctx.gate_batch("NAND", ct[:c], ct[:c])
    tm = ctx.collect_timing()
    br = tm["blind_rotate_ms"] / tm["blind_rotate_launches"]
    print("count %6d  BR %8.3f ms  per-592 %7.4f ms  gates/s %8.0f   KS %.3f ms" % (c, br, br / (c / 592), c / br * 1e3, tm["key_switch_ms"] / tm["key_switch_launches"]))
