"""GPU tests of the SURVEY 8(f) rank-4 breadth added in round 2 — each against the oracle:
  * LUT table + per-ciphertext index (tfhe_bootstrap_batch_indexed) == LUTs expanded;
  * many-LUT programmable bootstrap (tfhe_bootstrap_multi_lut_batch): k functions from one blind rotation;
  * opt-in two-blind-rotation MUX (tfhe_ctx_set_mux_mode) == the same composition of oracle primitives;
  * proxyreenc.ReencryptTLWELv0 (proxyreenc/proxyreenc.go:321-366) on the key-switch kernel, bit-exact."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    return importlib.import_module("go-tfhe_b200")


_CTX = {}


@pytest.fixture(scope="module")
def gpu(T, keyset):
    def get(name):
        if name not in _CTX:
            P, sk, ck = keyset(name)
            ctx = T.Context(T.params.get(name), 0)
            ctx.load_cloudkey(ck.offset, ck.bsk_fft, ck.ksk, ck.testvec)
            _CTX[name] = (P, sk, ck, ctx)
        return _CTX[name]
    yield get
    for v in _CTX.values():
        v[3].close()
    _CTX.clear()


@pytest.mark.parametrize("name,m", [("80", 2), ("uint3", 8)])
def test_indexed_luts_equal_expanded_luts(O, gpu, name, m):
    P, sk, ck, ctx = gpu(name)
    count = 23
    rng = np.random.default_rng(1)
    msgs = rng.integers(0, m, count)
    ct = sk.encrypt_message(msgs, m, 11)
    fs = [lambda x: x, lambda x: (m - 1) - x, lambda x: (x + 1) % m]
    luts = np.stack([O.gen_lut(P, m, f) for f in fs])
    idx = rng.integers(0, 3, count)
    got = ctx.bootstrap_batch_indexed(ct, luts, idx)
    assert np.array_equal(got, ctx.bootstrap_batch(ct, luts[idx]))
    assert list(sk.decrypt_message(got, m)) == [fs[k](int(v)) for k, v in zip(idx, msgs)]
    try:  # through the pipeline (index slices per chunk)
        ctx.set_pipeline_chunk(5)
        assert np.array_equal(ctx.bootstrap_batch_indexed(ct, luts, idx), got)
    finally:
        ctx.set_pipeline_chunk(16384)
    with pytest.raises(Exception):
        ctx.bootstrap_batch_indexed(ct, luts, np.full(count, 3))


def _switched(P, ct, lk):
    """The ciphertext whose REFERENCE mod switch (evaluator.go:116,122) gives the many-LUT mod switch of ct."""
    logn = int(np.log2(P.N))
    a = ct[:, :-1].astype(np.uint64)
    at = (((a + (1 << (30 - logn + lk))) % (1 << 32)) >> (31 - logn + lk)) << lk          # multiples of 2^lk in [0, 2N)
    out = np.empty_like(ct)
    out[:, :-1] = (at << (31 - logn)).astype(np.uint32)
    b = ct[:, -1].astype(np.uint64)
    bt = ((b + (1 << (30 - logn + lk))) >> (31 - logn + lk)) << lk                            # in [0, 2N], no wrap (int64 add)
    out[:, -1] = np.minimum(bt << (31 - logn), 0xFFFFFFFF).astype(np.uint32)
    return out


@pytest.mark.parametrize("name,m,lk", [("80", 2, 1), ("80", 2, 2), ("uint3", 8, 1)])
def test_many_lut_bootstrap(T, O, gpu, name, m, lk):
    """k = 2^lk functions per ciphertext from ONE blind rotation.  Decoded outputs exact; on the exact set every output
    word equals the oracle's blind rotation of the pre-switched ciphertext with the packed test vector, extracted at
    index i (trlwe.SampleExtractIndex) and key-switched."""
    P, sk, ck, ctx = gpu(name)
    k = 1 << lk
    msgs = np.arange(2 * m) % m
    ct = sk.encrypt_message(msgs, m, 21)
    fs = [lambda x: x, lambda x: (m - 1) - x, lambda x: (x + 1) % m, lambda x: (3 * x) % m][:k]
    packed = T.lut.PackLookUpTables([O.gen_lut(P, m, f) for f in fs]).reshape(1, -1)
    got = ctx.bootstrap_multi_lut_batch(ct, packed, lk)
    assert got.shape == (len(ct), k, P.n + 1)
    for i, f in enumerate(fs):
        assert list(sk.decrypt_message(got[:, i], m)) == [f(int(v)) for v in msgs], i
    if name == "80":
        ev = O.Evaluator(P.N)
        pre = _switched(P, ct, lk)
        for g in range(len(ct)):
            rot = ev.blind_rotate(P, pre[g], packed[0], ck.bsk_fft, ck.offset)
            for i in range(k):
                want = O.key_switch(P, O.sample_extract_index(rot, P.N, i), ck.ksk)
                assert np.array_equal(got[g, i], want), (g, i)
    # a batch past the latency kernel's range goes through the throughput kernel: same words
    big = np.concatenate([ct] * 80)[:300]
    assert np.array_equal(ctx.bootstrap_multi_lut_batch(big, packed, lk)[: len(ct)], got)


def test_two_blind_rotation_mux_is_opt_in_and_exact(O, gpu):
    """mode 1: MUX = KeySwitch(SampleExtract(BR(AND(a,b))) + SampleExtract(BR(ANDNY(a,c))) + 1/8).  Truth table, and the
    words of that very composition built from oracle primitives; mode 0 stays the reference's three-bootstrap MUX."""
    P, sk, ck, ctx = gpu("80")
    A, B, C = [0, 0, 0, 0, 1, 1, 1, 1, 1], [0, 0, 1, 1, 0, 0, 1, 1, 0], [0, 1, 0, 1, 0, 1, 0, 1, 1]
    a, b, c = sk.encrypt_bool(A, 51), sk.encrypt_bool(B, 52), sk.encrypt_bool(C, 53)
    ops = ["MUX"] * 8 + ["XOR"]
    ref = ctx.gate_batch(ops, a, b, c)
    assert np.array_equal(ref[:8], O.mux(ck, a[:8], b[:8], c[:8]))
    try:
        ctx.set_mux_mode(1)
        got = ctx.gate_batch(ops, a, b, c)
    finally:
        ctx.set_mux_mode(0)
    want_bits = [y if x else z for x, y, z in zip(A[:8], B[:8], C[:8])] + [A[8] ^ B[8]]
    assert list(sk.decrypt_bool(got)) == want_bits
    assert np.array_equal(got[8], ref[8])                 # non-MUX gates are untouched by the mode
    assert not np.array_equal(got[:8], ref[:8])           # different (valid) ciphertexts: why the mode is opt-in
    ev = O.Evaluator(P.N)
    for g in range(8):
        u1 = O.sample_extract0(ev.blind_rotate(P, O.gate_prepare(P, "AND", a[g], b[g]), ck.testvec, ck.bsk_fft, ck.offset), P.N)
        u2 = O.sample_extract0(ev.blind_rotate(P, O.gate_prepare(P, "ANDNY", a[g], c[g]), ck.testvec, ck.bsk_fft, ck.offset), P.N)
        s = (u1 + u2).astype(np.uint32)
        s[P.N] = np.uint32((int(s[P.N]) + 0x20000000) & 0xFFFFFFFF)
        assert np.array_equal(got[g], O.key_switch(P, s, ck.ksk)), g
    assert np.array_equal(ctx.gate_batch(ops, a, b, c), ref)   # back to the reference's MUX


def test_proxy_reencryption_bit_exact(T, O, gpu):
    """proxyreenc.ReencryptTLWELv0 == the engine's key-switch kernel with source dimension n.  The key is built as
    NewProxyReencryptionKeySymmetric does (proxyreenc.go:249-300): row (base t i + base j + k) encrypts
    k * keyFrom[i] / 2^((j+1) basebit) under keyTo; k = 0 rows stay zero."""
    P, sk, ck, ctx = gpu("80")
    n, basebit, t = P.n, P.basebit, P.iks_t
    base = 1 << basebit
    sk_to = O.SecretKey(P, 4242)
    rng = np.random.default_rng(7)
    rows = n * t * base
    key = np.zeros((rows, n + 1), dtype=np.uint32)
    i, j, k = np.meshgrid(np.arange(n), np.arange(t), np.arange(base), indexing="ij")
    mu = (k * sk.s0[i].astype(np.float64)) / (2.0 ** ((j + 1) * basebit))
    mask = rng.integers(0, 1 << 32, (rows, n), dtype=np.uint64).astype(np.uint32)
    noise = np.rint(rng.normal(0.0, P.alpha_lv0 * 2.0 ** 32, rows)).astype(np.int64)
    bterm = (mask.astype(np.uint64) * sk_to.s0.astype(np.uint64)).sum(1) + (np.fmod(mu.ravel(), 1.0) * 2.0 ** 32).astype(np.int64).astype(np.uint64) + noise.astype(np.uint64)
    key[:, :n] = mask
    key[:, n] = (bterm % (1 << 32)).astype(np.uint32)
    key[(k == 0).ravel()] = 0
    bits = np.array([0, 1, 1, 0, 1, 0, 0, 1, 1, 1], dtype=np.uint8)
    ct = sk.encrypt_bool(bits, 61)
    ctx.load_reencryption_key(key, basebit, t)
    got = ctx.reencrypt_batch(ct)
    want = np.stack([O.reencrypt(P, c, key, basebit, t) for c in ct])
    assert np.array_equal(got, want)
    assert list(sk_to.decrypt_bool(got)) == list(bits)           # same plaintexts, now under the target key
    big = np.concatenate([ct] * 40)                                 # past the split-gather range, through the pipeline
    try:
        ctx.set_pipeline_chunk(150)
        assert np.array_equal(ctx.reencrypt_batch(big)[:10], want)
    finally:
        ctx.set_pipeline_chunk(16384)
