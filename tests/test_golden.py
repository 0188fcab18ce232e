"""Committed golden fixtures (tests/golden/bootstrap_golden.json, made by tests/golden/make_golden.py from the oracle
with fixed seeds): the oracle must still reproduce them bit for bit (CPU), and so must the CUDA engine (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "bootstrap_golden.json")))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint32).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def material(O, keyset):
    P, sk, ck = keyset(GOLD["params"])
    a = sk.encrypt_bool([0, 0, 1, 1], GOLD["input_seeds"][0])
    b = sk.encrypt_bool([0, 1, 0, 1], GOLD["input_seeds"][1])
    assert [digest(a), digest(b)] == GOLD["inputs_sha256"]
    assert digest(ck.ksk[:4096]) == GOLD["ksk_sha256"]
    assert hashlib.sha256(ck.bsk_fft[0].tobytes()).hexdigest() == GOLD["bsk_row0_sha256"]
    return P, sk, ck, a, b


def test_oracle_reproduces_golden_gates(O, material):
    P, sk, ck, a, b = material
    for op, g in GOLD["gates"].items():
        r = O.gate_batch(ck, op, a, b)
        assert [int(x) for x in r[:, :4].ravel()] == g["head"], op
        assert digest(r) == g["sha256"], op
        assert [int(x) for x in sk.decrypt_bool(r)] == g["decrypted"], op


def test_oracle_reproduces_golden_blind_rotate_and_pbs(O, material):
    P, sk, ck, a, b = material
    rot = O.Evaluator(P.N).blind_rotate(P, a[3], ck.testvec, ck.bsk_fft, ck.offset)
    assert digest(rot) == GOLD["blind_rotate"]["sha256"]
    lut = O.gen_lut(P, 2, lambda x: 1 - x)
    assert digest(lut) == GOLD["pbs_not"]["lut_sha256"]
    pbs = O.bootstrap_batch(ck, sk.encrypt_message([0, 1], 2, 9003), lut)
    assert digest(pbs) == GOLD["pbs_not"]["sha256"]
    assert [int(x) for x in sk.decrypt_message(pbs, 2)] == GOLD["pbs_not"]["decoded"] == [1, 0]


@pytest.mark.gpu
def test_cuda_engine_reproduces_golden(material):
    import importlib
    T = importlib.import_module("go-tfhe_b200")
    P, sk, ck, a, b = material
    ctx = T.Context(T.params.get(GOLD["params"]), 0)
    try:
        ctx.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
        for op, g in GOLD["gates"].items():
            r = ctx.gate_batch(op, a, b)
            assert [int(x) for x in r[:, :4].ravel()] == g["head"], op
            assert digest(r) == g["sha256"], op
        rot = ctx.blind_rotate_batch(a[3:4]).reshape(-1)
        assert digest(rot) == GOLD["blind_rotate"]["sha256"]
        lut = np.asarray(__import__("oracle.oracle", fromlist=["x"]).gen_lut(P, 2, lambda x: 1 - x))
        pbs = ctx.bootstrap_batch(sk.encrypt_message([0, 1], 2, 9003), lut)
        assert digest(pbs) == GOLD["pbs_not"]["sha256"]
    finally:
        ctx.close()
