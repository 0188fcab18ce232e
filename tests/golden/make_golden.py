"""Generates tests/golden/bootstrap_golden.json from the CPU oracle with fixed seeds.

The reference cannot run in this image (pure Go, no toolchain) and ships no ciphertext vectors, so these fixtures
do NOT pin the oracle to the reference; they pin (a) the oracle against silent drift and (b) the CUDA path against
committed bits: tests/test_golden.py recomputes the same quantities with the oracle (CPU) and with the GPU engine
(-m gpu) and compares SHA-256 digests plus the leading words stored here.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OPS = ["NAND", "AND", "OR", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint32).tobytes()).hexdigest()


def build(name="80", sk_seed=0xC0FFEE + 2, ck_seed=0xBEEF):
    """Same seeds as tests/conftest.py::keyset so the session key set can be reused."""
    P = O.get_params(name)
    sk = O.SecretKey(P, sk_seed)
    ck = O.CloudKey(sk, ck_seed)
    a = sk.encrypt_bool([0, 0, 1, 1], 9001)
    b = sk.encrypt_bool([0, 1, 0, 1], 9002)
    out = {"params": name, "sk_seed": sk_seed, "ck_seed": ck_seed, "input_seeds": [9001, 9002],
           "inputs_sha256": [digest(a), digest(b)], "ksk_sha256": digest(ck.ksk[:4096]),
           "bsk_row0_sha256": hashlib.sha256(ck.bsk_fft[0].tobytes()).hexdigest(), "gates": {}}
    for op in OPS:
        r = O.gate_batch(ck, op, a, b)
        out["gates"][op] = {"sha256": digest(r), "head": [int(x) for x in r[:, :4].ravel()],
                            "decrypted": [int(x) for x in sk.decrypt_bool(r)]}
    ev = O.Evaluator(P.N)
    rot = ev.blind_rotate(P, a[3], ck.testvec, ck.bsk_fft, ck.offset)
    out["blind_rotate"] = {"sha256": digest(rot), "head": [int(x) for x in rot[:8]]}
    msgs = sk.encrypt_message([0, 1], 2, 9003)
    lut = O.gen_lut(P, 2, lambda x: 1 - x)
    pbs = O.bootstrap_batch(ck, msgs, lut)
    out["pbs_not"] = {"sha256": digest(pbs), "lut_sha256": digest(lut), "decoded": [int(x) for x in sk.decrypt_message(pbs, 2)]}
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bootstrap_golden.json")
    with open(path, "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote", path)
