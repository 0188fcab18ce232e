"""Pins the CPU oracle against every known-answer fact the reference's own tests hold for the
bootstrap path.  The reference has no ciphertext-level vectors (unseeded RNG),
so these are plaintext-level pins plus independent exact-arithmetic checks."""
import numpy as np
import pytest

TRUTH = {  # gates/gates_test.go:27-34,53-60,79-86,105-112,131-138,157-164,183-190,209-216,235-242,261-268
    "NAND": [1, 1, 1, 0], "AND": [0, 0, 0, 1], "OR": [0, 1, 1, 1], "XOR": [0, 1, 1, 0], "XNOR": [1, 0, 0, 1],
    "NOR": [1, 0, 0, 0], "ANDNY": [0, 1, 0, 0], "ANDYN": [0, 0, 1, 0], "ORNY": [1, 1, 0, 1], "ORYN": [1, 0, 1, 1],
}


def negacyclic_exact(a, b):
    """Exact product in Z[X]/(X^N+1) with Python big integers, reduced mod 2^32."""
    N = len(a)
    a = [int(x) for x in a]
    b = [int(x) for x in b]
    out = [0] * N
    for i, ai in enumerate(a):
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            k = i + j
            if k < N:
                out[k] += ai * bj
            else:
                out[k - N] -= ai * bj
    return np.array([x % (1 << 32) for x in out], dtype=np.uint32)


def test_f64_to_torus_constants(O):  # utils/utils_test.go:15-19,32-33
    for d, want in [(0.0, 0), (0.125, 536870912), (-0.125, 3758096384), (0.25, 1073741824), (0.5, 2147483648)]:
        assert O.f64_to_torus(d) == want


def test_param_tables(O):  # params/params.go:83-391, params/params_test.go
    want = {"80": (550, 1024, 6, 3, 2, 7), "110": (630, 1024, 6, 3, 2, 8), "128": (700, 1024, 6, 3, 2, 9),
            "uint1": (700, 1024, 10, 2, 2, 8), "uint2": (687, 512, 18, 1, 4, 3), "uint3": (820, 1024, 23, 1, 6, 2),
            "uint4": (820, 2048, 22, 1, 5, 3), "uint5": (1071, 2048, 22, 1, 6, 3)}
    for name, w in want.items():
        P = O.get_params(name)
        assert (P.n, P.N, P.bgbit, P.L, P.basebit, P.iks_t) == w
        assert 1 << P.nbit == P.N
    assert O.lib().oracle_decomposition_offset(__import__("ctypes").byref(O.get_params("128"))) == 0x82080000
    assert O.lib().oracle_decomposition_offset(__import__("ctypes").byref(O.get_params("uint5"))) == 0x80000000


@pytest.mark.parametrize("N", [512, 1024, 2048])
def test_fft_round_trip(O, N):  # poly/poly_test.go:10-33 (tolerance 10 LSB there; observed 0 here)
    ev = O.Evaluator(N)
    p = (np.arange(N, dtype=np.uint64) * 12345).astype(np.uint32)
    q = ev.to_poly(ev.to_fourier(p))
    d = np.abs(q.astype(np.int64) - p.astype(np.int64))
    assert d.max() <= 10


def test_fourier_layout_is_evaluation_at_odd_roots(O):
    """FourierPoly (poly/poly.go:54-62) = values P(w^e), w = exp(i*pi/N), e = 1 mod 4, in groups of 4 re + 4 im."""
    N = 64
    ev = O.Evaluator(N)
    rng = np.random.default_rng(1)
    p = rng.integers(-100, 100, N).astype(np.int64)
    fp = ev.to_fourier(p.astype(np.uint32))
    vals = np.array([complex(fp[(k // 4) * 8 + k % 4], fp[(k // 4) * 8 + 4 + k % 4]) for k in range(N // 2)])
    roots = [np.exp(1j * np.pi * e / N) for e in range(1, 2 * N, 4)]
    evals = np.array([sum(int(c) * r ** j for j, c in enumerate(p)) for r in roots])
    for v in vals:  # every output is the evaluation at one of the e = 1 (mod 4) roots, each used once
        assert np.min(np.abs(evals - v)) < 1e-6
    assert len({int(np.argmin(np.abs(evals - v))) for v in vals}) == N // 2


@pytest.mark.parametrize("N", [512, 1024])
def test_mul_poly_is_exact_negacyclic_product(O, N):  # poly/poly_mul.go:12-22 as used by trlwe/trlwe.go:43
    ev = O.Evaluator(N)
    rng = np.random.default_rng(7)
    a = rng.integers(0, 1 << 32, N, dtype=np.uint64).astype(np.uint32)
    s = rng.integers(0, 2, N).astype(np.uint32)
    # MulPoly reads both operands as signed int32 (fourier_transform.go:73-83)
    a_signed = a.astype(np.int32).astype(np.int64)
    assert np.array_equal(ev.mul_poly(a, s), negacyclic_exact(a_signed, s))


def test_poly_mul_xk_quirk(O):  # poly/buffer_methods.go:133-164: wrap-around "negation" is 0xFFFFFFFF - a
    a = np.arange(1, 17, dtype=np.uint32)
    assert np.array_equal(O.poly_mul_xk(a, 0), a)
    assert np.array_equal(O.poly_mul_xk(a, 32), a)
    r = O.poly_mul_xk(a, 3)
    assert list(r[3:]) == list(a[:13]) and list(r[:3]) == [0xFFFFFFFF - int(x) for x in a[13:]]
    r = O.poly_mul_xk(a, 16 + 3)
    assert list(r[3:]) == [0xFFFFFFFF - int(x) for x in a[:13]] and list(r[:3]) == list(a[13:])


def test_decompose_reconstructs(O):  # poly/decomposer.go:55-66, cloudkey/cloudkey.go:60-71
    P = O.get_params("128")
    rng = np.random.default_rng(3)
    p = rng.integers(0, 1 << 32, P.N, dtype=np.uint64).astype(np.uint32)
    d = O.decompose(P, p, 0x82080000).astype(np.int32).astype(np.int64)
    assert d.min() >= -32 and d.max() < 32
    rec = sum(d[i] << (32 - (i + 1) * P.bgbit) for i in range(P.L))
    err = (rec - p.astype(np.int64) + (1 << 31)) % (1 << 32) - (1 << 31)
    # the offset has no rounding term (cloudkey.go:60-71), so the digits truncate: -2^(32-L*bgbit) < err <= 0
    assert err.max() <= 0 and err.min() > -(1 << (32 - P.L * P.bgbit))


def test_external_product_is_exact_at_n1024(O, keyset):
    """At N=1024, L=3, Bg=64 the f64-FFT external product equals the exact integer result."""
    P, sk, ck = keyset("80", with_ksk=False)
    ev = O.Evaluator(P.N)
    rng = np.random.default_rng(11)
    trlwe = rng.integers(0, 1 << 32, 2 * P.N, dtype=np.uint64).astype(np.uint32)
    row = ck.bsk_fft[5]
    got = ev.external_product(P, row, trlwe, ck.offset)
    # exact: recover the integer TRGSW rows from the Fourier key, then big-int convolution of the digits
    digs = np.concatenate([O.decompose(P, trlwe[:P.N], ck.offset), O.decompose(P, trlwe[P.N:], ck.offset)])
    digs = digs.astype(np.int32).astype(np.int64)
    accA = np.zeros(P.N, dtype=object)
    accB = np.zeros(P.N, dtype=object)
    for r in range(2 * P.L):
        ka = ev.to_poly(row[r, 0]).astype(np.int32).astype(np.int64)
        kb = ev.to_poly(row[r, 1]).astype(np.int32).astype(np.int64)
        accA = accA + negacyclic_exact(digs[r], ka).astype(object)
        accB = accB + negacyclic_exact(digs[r], kb).astype(object)
    want = np.concatenate([np.array([int(x) % (1 << 32) for x in accA], dtype=np.uint32),
                           np.array([int(x) % (1 << 32) for x in accB], dtype=np.uint32)])
    assert np.array_equal(got, want)


def test_sample_extract_and_keyswitch_decrypt(O, keyset):  # trlwe_ops.go:10-21, keyswitch.go:10-37
    P, sk, ck = keyset("80")
    ev = O.Evaluator(P.N)
    ct = sk.encrypt_bool([1], 5)[0]
    rot = ev.blind_rotate(P, ct, ck.testvec, ck.bsk_fft, ck.offset)
    ext = O.sample_extract0(rot, P.N)
    # phase of the extracted LWE under s1 is ~ +1/8
    ph = (int(ext[P.N]) - int(np.sum(ext[:P.N].astype(np.uint64) * sk.s1.astype(np.uint64)) % (1 << 32))) % (1 << 32)
    assert abs(ph - 0x20000000) < 1 << 28
    out = O.key_switch(P, ext, ck.ksk)
    assert abs(int(sk.phase(out)[0]) - 0x20000000) < 1 << 28


@pytest.mark.parametrize("name", ["80", "128"])
def test_gate_truth_tables(O, keyset, name):  # gates/gates_test.go:23-281 (128-bit default there)
    P, sk, ck = keyset(name)
    a = sk.encrypt_bool([0, 0, 1, 1], 21)
    b = sk.encrypt_bool([0, 1, 0, 1], 22)
    for op, want in TRUTH.items():
        assert list(sk.decrypt_bool(O.gate_batch(ck, op, a, b))) == want, op


def test_not_copy_constant(O, keyset):  # gates/gates_test.go:283-336
    P, sk, ck = keyset("80")
    a = sk.encrypt_bool([0, 1], 31)
    assert list(sk.decrypt_bool(O.NOT(a))) == [1, 0]
    assert list(sk.decrypt_bool(a.copy())) == [0, 1]
    assert sk.decrypt_bool(O.constant(P, True))[0] == 1 and sk.decrypt_bool(O.constant(P, False))[0] == 0
    assert int(O.constant(P, False)[P.n]) == 0xE0000001  # 1 - 0x20000000 in uint32 (gates.go:63-65)


def test_mux(O, keyset):  # gates/gates_test.go:338-366
    P, sk, ck = keyset("80")
    A = [0, 0, 0, 0, 1, 1, 1, 1]
    B = [0, 0, 1, 1, 0, 0, 1, 1]
    C = [0, 1, 0, 1, 0, 1, 0, 1]
    r = O.mux(ck, sk.encrypt_bool(A, 41), sk.encrypt_bool(B, 42), sk.encrypt_bool(C, 43))
    assert list(sk.decrypt_bool(r)) == [b if a else c for a, b, c in zip(A, B, C)]


def test_batch_and_or_xor(O, keyset):  # gates/gates_test.go:369-480, batch = 4
    P, sk, ck = keyset("80")
    a = sk.encrypt_bool([1, 1, 0, 0], 51)
    b = sk.encrypt_bool([1, 0, 1, 0], 52)
    assert list(sk.decrypt_bool(O.gate_batch(ck, "AND", a, b))) == [1, 0, 0, 0]
    assert list(sk.decrypt_bool(O.gate_batch(ck, "OR", a, b))) == [1, 1, 1, 0]
    assert list(sk.decrypt_bool(O.gate_batch(ck, "XOR", a, b))) == [0, 1, 1, 0]


def test_pbs_binary_80bit(O, keyset):  # evaluator/programmable_bootstrap_test.go:13-188
    P, sk, ck = keyset("80")
    ct = sk.encrypt_message([0, 1], 2, 61)
    ident = O.gen_lut(P, 2, lambda x: x)
    assert np.all(ident[:P.N] == 0)  # lut/debug_test.go:82 — A is all zero
    assert list(sk.decrypt_message(O.bootstrap_batch(ck, ct, ident), 2)) == [0, 1]
    assert list(sk.decrypt_message(O.bootstrap_batch(ck, ct, O.gen_lut(P, 2, lambda x: 1 - x)), 2)) == [1, 0]
    assert list(sk.decrypt_message(O.bootstrap_batch(ck, ct, O.gen_lut(P, 2, lambda x: 1)), 2)) == [1, 1]


def _uint_values(m):  # params/uint_params_test.go:131-147
    return list(range(m)) if m <= 8 else [0, 1, 2, m // 2, m - 3, m - 2, m - 1]


@pytest.mark.parametrize("name,m", [("uint2", 4), ("uint3", 8), ("uint5", 32)])
def test_pbs_uint_sets(O, keyset, name, m):  # params/uint_params_test.go:61-126
    P, sk, ck = keyset(name)
    xs = _uint_values(m)
    ct = sk.encrypt_message(xs, m, 71)
    assert list(sk.decrypt_message(ct, m)) == xs
    for f in (lambda x: x, lambda x: (m - 1) - x, lambda x: x % (m // 2)):
        got = sk.decrypt_message(O.bootstrap_batch(ck, ct, O.gen_lut(P, m, f)), m)
        assert list(got) == [f(x) for x in xs]


def test_key_switch_as_byte_plane_contraction(O, keyset):
    """The formulation behind the tensor-core key switch (go-tfhe_b200/csrc/key_switch_mma.cuh), modelled in numpy
    against the oracle's trgsw/keyswitch.go:10-37: out = (0,..,0,b) - S @ KSK with S the 0/1 selection of the non-zero
    digits (K = N*t*(base-1), kidx = (i*t + j)*(base-1) + (k-1)), evaluated as four u8 x u8 -> s32 products over the byte
    planes of the key and recombined with shifts mod 2^32.  Any summation order gives the same words."""
    P, sk, ck = keyset("80")
    base, t, n, N = P.base, P.iks_t, P.n, P.N
    rng = np.random.default_rng(21)
    ext = rng.integers(0, 1 << 32, (3, N + 1), dtype=np.uint64).astype(np.uint32)
    ext[1, :N] = 0
    K = N * t * (base - 1)
    ksk = ck.ksk.reshape(N, t, base, n + 1)
    rows = ksk[:, :, 1:, :].reshape(K, n + 1)                       # k = 0 rows dropped: kidx order
    planes = np.stack([(rows >> (8 * p)) & 0xFF for p in range(4)]).astype(np.int64)   # [4][K][n+1], values < 256
    prec = np.uint32(1 << (32 - (1 + P.basebit * t)))
    for g in range(len(ext)):
        abar = ext[g, :N] + prec
        S = np.zeros(K, dtype=np.int64)
        for j in range(t):
            k = (abar >> np.uint32(32 - (j + 1) * P.basebit)) & np.uint32(base - 1)
            i = np.nonzero(k)[0]
            S[(i * t + j) * (base - 1) + (k[i].astype(np.int64) - 1)] = 1
        assert S.sum() <= N * t
        D = np.einsum("k,pkw->pw", S, planes)                       # four exact integer products, each < 2^31
        assert D.max() < (1 << 31)
        total = sum(D[p] << (8 * p) for p in range(4)) & 0xFFFFFFFF
        out = (-total) & 0xFFFFFFFF
        out[n] = (out[n] + int(ext[g, N])) & 0xFFFFFFFF
        assert np.array_equal(out.astype(np.uint32), O.key_switch(P, ext[g], ck.ksk))
