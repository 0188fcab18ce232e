"""First-light timing: 128-bit NAND batch on one GPU with per-stage timing."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
T = importlib.import_module("go-tfhe_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "128"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
P = T.params.get(name)
sk = T.key.NewSecretKey(P, 1)
t = time.time(); ck = T.cloudkey.NewCloudKey(sk, 2); print("keygen s", time.time() - t, flush=True)
ctx = ck.engine(0)
rng = np.random.default_rng(0)
A = rng.integers(0, 2, count).astype(np.uint8); B = rng.integers(0, 2, count).astype(np.uint8)
a = T.tlwe.EncryptBool(A, sk, 3); b = T.tlwe.EncryptBool(B, sk, 4)
ctx.set_timing(True)
for it in range(3):
    t = time.time(); out = ctx.gate_batch("NAND", a, b); dt = time.time() - t
    tm = ctx.collect_timing()
    ok = np.array_equal(T.tlwe.DecryptBool(out, sk), 1 - (A & B))
    print("iter", it, "e2e s %.4f" % dt, "gates/s %.0f" % (count / dt), tm, "correct", ok, flush=True)
print("BR gates/s %.0f" % (count / (tm["blind_rotate_ms"] * 1e-3)), "KS gates/s %.0f" % (count / (tm["key_switch_ms"] * 1e-3)))
