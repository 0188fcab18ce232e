"""Device key generation (tfhe_ctx_generate_cloudkey) against the CPU oracle.

The reference seeds every RNG from unseeded math/rand (cloudkey/cloudkey.go:24-145, key/key.go:17), so there are no key
bytes to compare: the checks are (1) the exported key is a valid reference-format CloudKey — the ORACLE's own gate
evaluation with it decrypts to the truth tables; (2) the engine, keeping the same key resident, is bit-exact with the
oracle on that key; (3) every ciphertext of the key decrypts to its prescribed plaintext with noise of the prescribed
standard deviation (KSKAlpha / BSKAlpha); (4) determinism in (secret key, seed)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    return importlib.import_module("go-tfhe_b200")


class _OCK:
    pass


def _oracle_view(O, name, ck):
    o = _OCK()
    o.P, o.testvec, o.ksk, o.bsk_fft, o.offset = (O.get_params(name), ck.BlindRotateTestvec.ravel(), ck.KeySwitchingKey,
                                                  ck.BootstrappingKey, ck.DecompositionOffset)
    return o


@pytest.mark.parametrize("name", ["80", "uint2"])
def test_generated_key_is_a_valid_reference_cloudkey(T, O, name):
    P = T.params.get(name)
    OP = O.get_params(name)
    osk = O.SecretKey(OP, 77)
    sk = T.key.SecretKey(P, osk.s0.copy(), osk.s1.copy())
    ck = T.cloudkey.NewCloudKeyOnDevice(sk, seed=2024)
    try:
        ref = O.CloudKey(osk, 1, with_ksk=False, with_bsk=False)
        assert ck.DecompositionOffset == ref.offset                      # cloudkey.go:60-71
        assert np.array_equal(ck.BlindRotateTestvec.ravel(), ref.testvec)  # cloudkey.go:74-85
        ctx = ck.engine(0)
        ock = _oracle_view(O, name, ck)
        if name == "80":
            a = osk.encrypt_bool([0, 0, 1, 1], 5)
            b = osk.encrypt_bool([0, 1, 0, 1], 6)
            want = O.gate_batch(ock, "NAND", a, b)                       # the oracle computing with the GPU-made key
            assert list(osk.decrypt_bool(want)) == [1, 1, 1, 0]
            got = ctx.gate_batch("NAND", a, b)                           # the engine with the same key, still resident
            assert np.array_equal(got, want)
            got = ctx.gate_batch("XOR", a, b)
            assert list(osk.decrypt_bool(got)) == [0, 1, 1, 0]
        else:
            msgs = np.arange(4) % 4
            ct = osk.encrypt_message(msgs, 4, 9)
            lut = O.gen_lut(OP, 4, lambda x: (3 - x) % 4)
            got = ctx.bootstrap_batch(ct, np.asarray(lut).reshape(1, -1))
            assert list(osk.decrypt_message(got, 4)) == [3, 2, 1, 0]
            want = O.bootstrap_batch(ock, ct, np.asarray(lut).reshape(1, -1))
            assert list(osk.decrypt_message(want, 4)) == [3, 2, 1, 0]
    finally:
        ck.close()


def test_generated_key_noise_and_plaintexts(T, O):
    name = "80"
    P = T.params.get(name)
    OP = O.get_params(name)
    osk = O.SecretKey(OP, 78)
    sk = T.key.SecretKey(P, osk.s0.copy(), osk.s1.copy())
    ck = T.cloudkey.NewCloudKeyOnDevice(sk, seed=31337)
    try:
        base, t, n, N, L = 1 << P.BASEBIT, P.IKS_T, P.n, P.N, P.L
        ksk = ck.KeySwitchingKey.reshape(N, t, base, n + 1)
        # k = 0 rows stay zero (cloudkey.go:104-106); k >= 1 rows encrypt k * s1[i] / base^(j+1) under s0
        assert not ksk[:, :, 0, :].any()
        rows = ksk[:256, :, 1:, :].reshape(-1, n + 1)
        ph = osk.phase(rows).reshape(256, t, base - 1)
        i, j, k = np.meshgrid(np.arange(256), np.arange(t), np.arange(1, base), indexing="ij")
        mu = (k.astype(np.float64) * osk.s1[i]) / (2.0 ** ((j + 1) * P.BASEBIT))
        mu_t = (np.fmod(mu, 1.0) * 2.0 ** 32).astype(np.int64).astype(np.uint32)
        err = (ph - mu_t).astype(np.int32).astype(np.float64) / 2.0 ** 32
        assert abs(err.mean()) < 4 * P.alpha_lv0 / np.sqrt(err.size) + 2.0 ** -32
        assert 0.93 * P.alpha_lv0 < err.std() < 1.07 * P.alpha_lv0
        # masks are uniform: every bit of the mask words is balanced
        bits = np.unpackbits(rows[:, :n].view(np.uint8))
        assert abs(bits.mean() - 0.5) < 1e-3
        # bootstrapping key: inverse transform of every row of a few steps gives a TRLWE of 0 plus the gadget
        ev = O.Evaluator(N)
        errs = []
        for step in (0, 1, n // 2, n - 1):
            for r in range(2 * L):
                A = ev.to_poly(ck.BootstrappingKey[step, r, 0])
                B = ev.to_poly(ck.BootstrappingKey[step, r, 1])
                lvl = r if r < L else r - L
                g = (int(osk.s0[step]) << (32 - (lvl + 1) * P.BGBIT)) & 0xFFFFFFFF
                A, B = A.copy(), B.copy()
                if r < L:   # take the gadget term off again (trgsw.go:50-54) to recover the fresh encryption of zero
                    A[0] = np.uint32((int(A[0]) - g) & 0xFFFFFFFF)
                else:
                    B[0] = np.uint32((int(B[0]) - g) & 0xFFFFFFFF)
                e = (B - ev.mul_poly(A, osk.s1)).astype(np.int32).astype(np.float64) / 2.0 ** 32
                errs.append(e)
        errs = np.concatenate(errs)
        assert 0.9 * P.alpha_lv1 < errs.std() < 1.1 * P.alpha_lv1
        assert np.abs(errs).max() < 8 * P.alpha_lv1
    finally:
        ck.close()


def test_generated_key_is_deterministic_in_seed(T, O):
    P = T.params.get("uint2")
    sk = T.key.NewSecretKey(P, 5)
    a = T.cloudkey.NewCloudKeyOnDevice(sk, seed=1)
    b = T.cloudkey.NewCloudKeyOnDevice(sk, seed=1)
    c = T.cloudkey.NewCloudKeyOnDevice(sk, seed=2)
    try:
        assert np.array_equal(a.BootstrappingKey, b.BootstrappingKey) and np.array_equal(a.KeySwitchingKey, b.KeySwitchingKey)
        assert not np.array_equal(a.BootstrappingKey, c.BootstrappingKey)
        assert not np.array_equal(a.KeySwitchingKey, c.KeySwitchingKey)
    finally:
        a.close(); b.close(); c.close()


def test_device_keygen_uint5_without_export(T):
    """N = 2048: the key-generation kernel needs more than the default 48 KiB of dynamic shared memory (exchange buffers
    + the row's ChaCha20 mask words); the 1.57 GiB key-switching key is generated, repacked and used without ever
    leaving the device.  Decoded programmable bootstraps must be exact."""
    P = T.params.get("uint5")
    sk = T.key.NewSecretKey(P, 41)
    ctx = T.Context(P, 0)
    try:
        ctx.generate_cloudkey(sk.KeyLv0, sk.KeyLv1, seed=42, with_ksk=True, export=False)
        msgs = np.arange(32)
        ct = T.tlwe.EncryptLWEMessage(msgs, 32, sk, 43)
        lut = T.lut.NewGenerator(32, P).GenLookUpTable(lambda v: (v * 3 + 1) % 32).Poly.reshape(1, -1)
        got = ctx.bootstrap_batch(ct, lut)
        assert list(T.tlwe.DecryptLWEMessage(got, 32, sk)) == [(int(v) * 3 + 1) % 32 for v in msgs]
    finally:
        ctx.close()
