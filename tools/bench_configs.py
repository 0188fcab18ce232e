"""Secondary BASELINE.json configs on one GPU (configs[3], configs[4] per-GPU share, and configs[2] once the circuit
runner exists).  Prints one JSON line per config.  Not the driver's bench (that is bench.py = configs[1])."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
T = importlib.import_module("go-tfhe_b200")
which = sys.argv[1:] or ["c4", "c5", "c3"]


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


if "c4" in which:  # programmable bootstrap, Uint5 (n=1071, N=2048, msgMod=32), batch 2048
    P = T.params.get("uint5")
    sk = T.key.NewSecretKey(P, 1); ck = T.cloudkey.NewCloudKey(sk, 2); ctx = ck.engine(0)
    rng = np.random.default_rng(0); count = 2048
    msgs = rng.integers(0, 32, count)
    ct = T.tlwe.EncryptLWEMessage(msgs, 32, sk, 3)
    fs = [lambda x: x, lambda x: x % 16, lambda x: int(x >= 16)]
    luts = np.stack([T.lut.NewGenerator(32, P).GenLookUpTable(f).Poly for f in fs])
    sel = rng.integers(0, 3, count)
    per = luts[sel].reshape(count, -1)
    ctx.set_timing(True)
    dt = timed(lambda: ctx.bootstrap_batch(ct, per)); tm = ctx.collect_timing()
    out = ctx.bootstrap_batch(ct, per)
    dec = T.tlwe.DecryptLWEMessage(out, 32, sk)
    want = np.array([fs[s](int(m)) for s, m in zip(sel, msgs)])
    print(json.dumps({"config": "c4 PBS Uint5 batch 2048", "bootstraps_per_s_e2e": count / dt, "correct": bool(np.array_equal(dec, want)),
                      "blind_rotate_ms": tm["blind_rotate_ms"] / max(tm["blind_rotate_launches"], 1),
                      "key_switch_ms": tm["key_switch_ms"] / max(tm["key_switch_launches"], 1)}), flush=True)
    ck.close()

if "c5" in which or "c3" in which:
    P = T.params.get("128")
    sk = T.key.NewSecretKey(P, 1); ck = T.cloudkey.NewCloudKey(sk, 2); ctx = ck.engine(0)

if "c5" in which:  # 2^20 mixed AND/OR/XOR/MUX gates over 8 GPUs -> this GPU's share: 2^17 gate-ops
    rng = np.random.default_rng(1); count = 1 << 17; pool = 4096
    bits = rng.integers(0, 2, pool).astype(np.uint8)
    cts = T.tlwe.EncryptBool(bits, sk, 5)
    ia, ib, ic = (rng.integers(0, pool, count) for _ in range(3))
    ops = rng.integers(0, 4, count)
    opcodes = np.array([T.OPCODES[o] for o in ("AND", "OR", "XOR", "MUX")], dtype=np.uint8)[ops]
    a, b, c = cts[ia], cts[ib], cts[ic]
    dt = timed(lambda: ctx.gate_batch(opcodes, a, b, c), reps=1)
    out = ctx.gate_batch(opcodes, a, b, c)
    A, B, C = bits[ia], bits[ib], bits[ic]
    want = np.select([ops == 0, ops == 1, ops == 2], [A & B, A | B, A ^ B], np.where(A == 1, B, C))
    nboot = int(count + 2 * (ops == 3).sum())
    print(json.dumps({"config": "c5 mixed AND/OR/XOR/MUX, 2^17 gate-ops (1/8 of 2^20) on one GPU", "gate_ops_per_s_e2e": count / dt,
                      "bootstraps_per_s_e2e": nboot / dt, "bootstraps": nboot,
                      "correct": bool(np.array_equal(T.tlwe.DecryptBool(out, sk), want))}), flush=True)

if "c3" in which and hasattr(T, "circuit"):  # 8-bit ripple-carry adder x 1024 instances
    inst = 1024
    rng = np.random.default_rng(2)
    x, y = rng.integers(0, 256, inst), rng.integers(0, 256, inst)
    circ = T.circuit.ripple_carry_adder(8)
    xin = [T.tlwe.EncryptBool((x >> i) & 1, sk, 100 + i) for i in range(8)]
    yin = [T.tlwe.EncryptBool((y >> i) & 1, sk, 200 + i) for i in range(8)]
    inputs = np.stack(xin + yin)
    dt = timed(lambda: circ.run(ck, inputs), reps=2)
    outs = circ.run(ck, inputs)
    s = sum(T.tlwe.DecryptBool(outs[i], sk).astype(np.int64) << i for i in range(8))
    print(json.dumps({"config": "c3 8-bit ripple-carry adder x1024 (40 bootstraps each)", "adders_per_s_e2e": inst / dt,
                      "bootstraps_per_s_e2e": inst * 40 / dt, "levels": circ.n_levels,
                      "correct": bool(np.array_equal(s, (x + y) % 256))}), flush=True)
