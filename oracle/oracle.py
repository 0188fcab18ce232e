"""ctypes binding of oracle/liboracle.so — the CPU restatement of the go-tfhe bootstrap path.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py — never by go-tfhe_b200/.
Flattened layouts are identical to the C ABI in include/tfhe_b200.h:
  ciphertext  [count][n+1] u32         testvec / LUT  [2][N] u32 (A then B)
  ksk         [N][t][base][n+1] u32    bsk_fft        [n][2L][2][N] f64 (reference FourierPoly layout)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

OPS = {"NAND": 0, "AND": 1, "OR": 2, "XOR": 3, "XNOR": 4, "NOR": 5, "ANDNY": 6, "ANDYN": 7, "ORNY": 8, "ORYN": 9}


class Params(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("N", ctypes.c_int32), ("nbit", ctypes.c_int32), ("bgbit", ctypes.c_int32),
                ("L", ctypes.c_int32), ("basebit", ctypes.c_int32), ("iks_t", ctypes.c_int32), ("_pad", ctypes.c_int32),
                ("alpha_lv0", ctypes.c_double), ("alpha_lv1", ctypes.c_double)]

    @property
    def base(self):
        return 1 << self.basebit

    @property
    def ksk_rows(self):
        return self.N * self.iks_t * self.base

    @property
    def bsk_row_doubles(self):
        return 2 * self.L * 2 * self.N


def build():
    """Compile liboracle.so (g++ -O2 -ffp-contract=off) if missing or stale."""
    src = os.path.join(_HERE, "oracle.cpp")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_f64_to_torus.restype = ctypes.c_uint32
        _lib.oracle_f64_to_torus.argtypes = [ctypes.c_double]
        _lib.oracle_decomposition_offset.restype = ctypes.c_uint32
        _lib.oracle_eval_new.restype = ctypes.c_void_p
        _lib.oracle_eval_new.argtypes = [ctypes.c_int]
        _lib.oracle_eval_free.argtypes = [ctypes.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def get_params(name):
    p = Params()
    if lib().oracle_get_params(str(name).encode(), ctypes.byref(p)) != 0:
        raise KeyError(name)
    return p


def f64_to_torus(d):
    return int(lib().oracle_f64_to_torus(float(d)))


class Evaluator:
    """poly.Evaluator restatement: twiddles + scratch (poly/poly_evaluator.go:76)."""

    def __init__(self, N):
        self.N = N
        self.h = ctypes.c_void_p(lib().oracle_eval_new(N))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_eval_free(self.h)
            self.h = None

    def twiddles(self):
        cnt = lib().oracle_twiddle_count(self.h)
        tw = np.zeros(2 * cnt)
        twi = np.zeros(2 * cnt)
        lib().oracle_get_twiddles(self.h, _p(tw), _p(twi))
        return tw.view(np.complex128), twi.view(np.complex128)

    def to_fourier(self, p):
        p = np.ascontiguousarray(p, dtype=np.uint32)
        fp = np.zeros(self.N)
        lib().oracle_to_fourier(self.h, _p(p), _p(fp))
        return fp

    def to_poly(self, fp):
        fp = np.ascontiguousarray(fp, dtype=np.float64)
        out = np.zeros(self.N, dtype=np.uint32)
        lib().oracle_to_poly(self.h, _p(fp), _p(out))
        return out

    def mul_poly(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        out = np.zeros(self.N, dtype=np.uint32)
        lib().oracle_mul_poly(self.h, _p(a), _p(b), _p(out))
        return out

    def external_product(self, P, bsk_row, trlwe, offset):
        out = np.zeros(2 * P.N, dtype=np.uint32)
        lib().oracle_external_product(self.h, ctypes.byref(P), _p(np.ascontiguousarray(bsk_row)),
                                      _p(np.ascontiguousarray(trlwe, dtype=np.uint32)), ctypes.c_uint32(offset), _p(out))
        return out

    def cmux(self, P, bsk_row, ct0, ct1, offset):
        out = np.zeros(2 * P.N, dtype=np.uint32)
        lib().oracle_cmux(self.h, ctypes.byref(P), _p(np.ascontiguousarray(bsk_row)),
                          _p(np.ascontiguousarray(ct0, dtype=np.uint32)), _p(np.ascontiguousarray(ct1, dtype=np.uint32)),
                          ctypes.c_uint32(offset), _p(out))
        return out

    def blind_rotate(self, P, ct, testvec, bsk, offset):
        out = np.zeros(2 * P.N, dtype=np.uint32)
        lib().oracle_blind_rotate(self.h, ctypes.byref(P), _p(np.ascontiguousarray(ct, dtype=np.uint32)),
                                  _p(np.ascontiguousarray(testvec, dtype=np.uint32)), _p(bsk), ctypes.c_uint32(offset),
                                  _p(out))
        return out


def poly_mul_xk(a, k):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    out = np.zeros_like(a)
    lib().oracle_poly_mul_xk(_p(a), ctypes.c_int(len(a)), ctypes.c_int64(k), _p(out))
    return out


def decompose(P, p, offset):
    p = np.ascontiguousarray(p, dtype=np.uint32)
    out = np.zeros((P.L, P.N), dtype=np.uint32)
    lib().oracle_decompose(ctypes.byref(P), _p(p), ctypes.c_uint32(offset), _p(out))
    return out


def sample_extract0(trlwe, N):
    out = np.zeros(N + 1, dtype=np.uint32)
    lib().oracle_sample_extract0(_p(np.ascontiguousarray(trlwe, dtype=np.uint32)), ctypes.c_int(N), _p(out))
    return out


def key_switch(P, src, ksk):
    out = np.zeros(P.n + 1, dtype=np.uint32)
    lib().oracle_key_switch(ctypes.byref(P), _p(np.ascontiguousarray(src, dtype=np.uint32)), _p(ksk), _p(out))
    return out


def reencrypt(P, ct_from, key, basebit, t):
    """proxyreenc.ReencryptTLWELv0 (proxyreenc/proxyreenc.go:321-366); key = KeyEncryptions [n*t*base][n+1]."""
    out = np.zeros(P.n + 1, dtype=np.uint32)
    lib().oracle_reencrypt(ctypes.byref(P), _p(np.ascontiguousarray(ct_from, dtype=np.uint32)),
                           _p(np.ascontiguousarray(key, dtype=np.uint32)), ctypes.c_int(basebit), ctypes.c_int(t), _p(out))
    return out


def sample_extract_index(trlwe, N, k):
    """trlwe.SampleExtractIndex (trlwe/trlwe.go:114-128): out[i] = A[k-i] (i <= k), 0xFFFFFFFF - A[N+k-i] (i > k), out[N] = B[k]."""
    t = np.ascontiguousarray(trlwe, dtype=np.uint32).ravel()
    A, B = t[:N], t[N:]
    out = np.zeros(N + 1, dtype=np.uint32)
    i = np.arange(N)
    out[:N] = np.where(i <= k, A[(k - i) % N], np.uint32(0xFFFFFFFF) - A[(N + k - i) % N])
    out[N] = B[k]
    return out


def gate_prepare(P, op, a, b):
    out = np.zeros(P.n + 1, dtype=np.uint32)
    rc = lib().oracle_gate_prepare(ctypes.byref(P), ctypes.c_int(OPS[op] if isinstance(op, str) else op),
                                   _p(np.ascontiguousarray(a, dtype=np.uint32)),
                                   _p(np.ascontiguousarray(b, dtype=np.uint32)), _p(out))
    if rc:
        raise ValueError(op)
    return out


class SecretKey:
    """key.SecretKey (key/key.go:10-13) from a seed."""

    def __init__(self, P, seed):
        self.P = P
        self.s0 = np.zeros(P.n, dtype=np.uint32)
        self.s1 = np.zeros(P.N, dtype=np.uint32)
        lib().oracle_secret_key(ctypes.byref(P), ctypes.c_uint64(seed), _p(self.s0), _p(self.s1))

    def encrypt_bool(self, bits, seed):
        bits = np.ascontiguousarray(bits, dtype=np.uint8).ravel()
        out = np.zeros((len(bits), self.P.n + 1), dtype=np.uint32)
        lib().oracle_encrypt_bool(ctypes.byref(self.P), _p(self.s0), ctypes.c_uint64(seed), ctypes.c_int(len(bits)),
                                  _p(bits), _p(out))
        return out

    def decrypt_bool(self, ct):
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, self.P.n + 1)
        bits = np.zeros(len(ct), dtype=np.uint8)
        lib().oracle_decrypt_bool(ctypes.byref(self.P), _p(self.s0), ctypes.c_int(len(ct)), _p(ct), _p(bits))
        return bits

    def phase(self, ct):
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, self.P.n + 1)
        ph = np.zeros(len(ct), dtype=np.uint32)
        lib().oracle_phase(ctypes.byref(self.P), _p(self.s0), ctypes.c_int(len(ct)), _p(ct), _p(ph))
        return ph

    def encrypt_message(self, msgs, msg_mod, seed):
        msgs = np.ascontiguousarray(msgs, dtype=np.int32).ravel()
        out = np.zeros((len(msgs), self.P.n + 1), dtype=np.uint32)
        lib().oracle_encrypt_message(ctypes.byref(self.P), _p(self.s0), ctypes.c_uint64(seed), ctypes.c_int(len(msgs)),
                                     _p(msgs), ctypes.c_int(msg_mod), _p(out))
        return out

    def decrypt_message(self, ct, msg_mod):
        ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, self.P.n + 1)
        out = np.zeros(len(ct), dtype=np.int32)
        lib().oracle_decrypt_message(ctypes.byref(self.P), _p(self.s0), ctypes.c_int(len(ct)), _p(ct),
                                     ctypes.c_int(msg_mod), _p(out))
        return out


class CloudKey:
    """cloudkey.CloudKey (cloudkey/cloudkey.go:16-21), flattened."""

    def __init__(self, sk, seed, threads=None, with_ksk=True, with_bsk=True):
        P = sk.P
        self.P = P
        threads = threads or os.cpu_count() or 1
        off = ctypes.c_uint32(0)
        self.testvec = np.zeros(2 * P.N, dtype=np.uint32)
        self.ksk = np.zeros((P.ksk_rows, P.n + 1), dtype=np.uint32) if with_ksk else None
        self.bsk_fft = np.zeros((P.n, 2 * P.L, 2, P.N), dtype=np.float64) if with_bsk else None
        lib().oracle_cloudkey(ctypes.byref(P), _p(sk.s0), _p(sk.s1), ctypes.c_uint64(seed), ctypes.c_int(threads),
                              ctypes.byref(off), _p(self.testvec), _p(self.ksk), _p(self.bsk_fft))
        self.offset = int(off.value)


def gen_lut(P, msg_mod, f):
    """lut.Generator.GenLookUpTable (lut/generator.go:49-100): f is a callable or a table."""
    fv = np.array([f(x) for x in range(msg_mod)] if callable(f) else list(f), dtype=np.int32)
    out = np.zeros(2 * P.N, dtype=np.uint32)
    lib().oracle_gen_lut(ctypes.byref(P), ctypes.c_int(msg_mod), _p(fv), _p(out))
    return out


def bootstrap_batch(ck, ct_in, luts=None, threads=None):
    P = ck.P
    ct_in = np.ascontiguousarray(ct_in, dtype=np.uint32).reshape(-1, P.n + 1)
    out = np.zeros_like(ct_in)
    nluts = 0
    if luts is not None:
        luts = np.ascontiguousarray(luts, dtype=np.uint32).reshape(-1, 2 * P.N)
        nluts = len(luts)
    lib().oracle_bootstrap_batch(ctypes.byref(P), ctypes.c_int(len(ct_in)), _p(ct_in), _p(ck.testvec), _p(luts),
                                 ctypes.c_int(nluts), _p(ck.bsk_fft), _p(ck.ksk), ctypes.c_uint32(ck.offset),
                                 ctypes.c_int(threads or os.cpu_count() or 1), _p(out))
    return out


def gate_batch(ck, ops, a, b, threads=None):
    P = ck.P
    a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, P.n + 1)
    b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, P.n + 1)
    if isinstance(ops, (str, int)):
        ops = [ops]
    opv = np.array([OPS[o] if isinstance(o, str) else o for o in ops], dtype=np.uint8)
    out = np.zeros_like(a)
    rc = lib().oracle_gate_batch(ctypes.byref(P), ctypes.c_int(len(a)), _p(opv), ctypes.c_int(len(opv)), _p(a), _p(b),
                                 _p(ck.testvec), _p(ck.bsk_fft), _p(ck.ksk), ctypes.c_uint32(ck.offset),
                                 ctypes.c_int(threads or os.cpu_count() or 1), _p(out))
    if rc:
        raise ValueError("bad opcode")
    return out


def NOT(a):
    """gates.NOT (gates/gates.go:117-119): 0 - P."""
    return (np.uint32(0) - np.asarray(a, dtype=np.uint32)).astype(np.uint32)


def constant(P, value):
    """gates.Constant (gates/gates.go:61-69)."""
    mu = f64_to_torus(0.125)
    if not value:
        mu = (1 - mu) & 0xFFFFFFFF
    out = np.zeros(P.n + 1, dtype=np.uint32)
    out[P.n] = mu
    return out


def mux(ck, a, b, c, threads=None):
    """gates.MUX (gates/gates.go:107-114): OR(AND(a,b), AND(NOT a, c))."""
    and_ab = gate_batch(ck, "AND", a, b, threads)
    and_nac = gate_batch(ck, "AND", NOT(a), c, threads)
    return gate_batch(ck, "OR", and_ab, and_nac, threads)
