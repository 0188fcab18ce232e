import importlib, sys, os, numpy as np
sys.path.insert(0, os.getcwd())
T = importlib.import_module("go-tfhe_b200")
P = T.params.get("80"); sk = T.key.NewSecretKey(P, 1); ck = T.cloudkey.NewCloudKeyOnDevice(sk, 2, export=False); ctx = ck.engine(0)
a = T.tlwe.EncryptBool([1], sk, 3); b = T.tlwe.EncryptBool([0], sk, 4)
ctx.set_blind_rotate_variant(sys.argv[1])
for _ in range(3): out = ctx.gate_batch("NAND", a, b)
print(T.tlwe.DecryptBool(out, sk))
