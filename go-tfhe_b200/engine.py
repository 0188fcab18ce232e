"""Context: owns one tfhe_ctx (one GPU).  Thin, typed wrapper over the C ABI; all compute happens in
libtfhe_b200.so.  numpy arrays in / out for the host-buffer entry points, raw device pointers (ints,
e.g. torch.Tensor.data_ptr()) for the *_device entry points."""
import ctypes

import numpy as np

from . import _native
from .params import ParamSet


def _seed(seed):
    """None -> 0 (the C side then keys ChaCha20 from the OS entropy source); integers are reproducible test seeds."""
    if seed is None:
        return 0
    seed = int(seed) & (2**64 - 1)
    return seed if seed != 0 else 0x9E3779B97F4A7C15

OPCODES = {"NAND": 0, "AND": 1, "OR": 2, "XOR": 3, "XNOR": 4, "NOR": 5, "ANDNY": 6, "ANDYN": 7, "ORNY": 8, "ORYN": 9,
           "MUX": 10, "NOT": 11, "COPY": 12}


class TfheError(RuntimeError):
    pass


def _u32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a if shape is None else a.reshape(shape)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _ops(ops, count):
    if isinstance(ops, (str, int)):
        ops = [ops]
    v = np.array([OPCODES[o.upper()] if isinstance(o, str) else int(o) for o in ops], dtype=np.uint8)
    if len(v) not in (1, count):
        raise ValueError("ops must have length 1 or count")
    return v


class Context:
    def __init__(self, P: ParamSet, device=0, devices=None):
        """device: one GPU (tfhe_ctx_create).  devices: a list of GPUs, or "all" — ONE context that shards every host-buffer
        batch call over them (tfhe_ctx_create_multi)."""
        self.P = P
        self.lib = _native.engine()
        self.h = ctypes.c_void_p()
        tp = _native.TfheParams(P.n, P.N, P.L, P.BGBIT, P.BASEBIT, P.IKS_T)
        if devices is None:
            rc = self.lib.tfhe_ctx_create(ctypes.byref(tp), int(device), ctypes.byref(self.h))
            what = "tfhe_ctx_create"
        else:
            what = "tfhe_ctx_create_multi"
            if isinstance(devices, str):
                rc = self.lib.tfhe_ctx_create_multi(ctypes.byref(tp), 0, None, ctypes.byref(self.h))
            else:
                arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
                rc = self.lib.tfhe_ctx_create_multi(ctypes.byref(tp), len(devices), ctypes.cast(arr, ctypes.c_void_p), ctypes.byref(self.h))
        if rc != 0:
            raise TfheError("%s: %s" % (what, self.lib.tfhe_last_error(None).decode()))
        self.device = device

    @property
    def device_count(self):
        return int(self.lib.tfhe_ctx_device_count(self.h))

    def set_pipeline_chunk(self, rows):
        """ciphertexts per chunk of the pipelined host-buffer calls (results do not depend on it)."""
        self._ck(self.lib.tfhe_ctx_set_pipeline_chunk(self.h, int(rows)), "tfhe_ctx_set_pipeline_chunk")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.tfhe_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise TfheError("%s failed (%d): %s" % (what, rc, self.lib.tfhe_last_error(self.h).decode()))

    # --- keys ---------------------------------------------------------------------------------------
    def load_cloudkey(self, offset, bsk_fft, ksk, testvec):
        P = self.P
        bsk = np.ascontiguousarray(bsk_fft, dtype=np.float64)
        assert bsk.size == P.n * 2 * P.L * 2 * P.N, "bsk_fft must be [n][2L][2][N]"
        k = None
        if ksk is not None:
            k = _u32(ksk)
            assert k.size == P.ksk_rows * (P.n + 1), "ksk must be [N][t][base][n+1]"
        tv = _u32(testvec)
        assert tv.size == 2 * P.N
        self._ck(self.lib.tfhe_ctx_load_cloudkey(self.h, ctypes.c_uint32(offset), _ptr(bsk), _ptr(k), _ptr(tv)),
                 "tfhe_ctx_load_cloudkey")

    def generate_cloudkey(self, key_lv0, key_lv1, seed=None, with_ksk=True, export=True):
        """cloudkey.NewCloudKey on the device (tfhe_ctx_generate_cloudkey).  Returns (offset, testvec, ksk, bsk_fft) in the
        reference layouts when export=True (ksk None without a key-switching key), else None; the key stays loaded."""
        P = self.P
        s0, s1 = _u32(key_lv0), _u32(key_lv1)
        assert s0.size == P.n and s1.size == P.N
        off = ctypes.c_uint32(0)
        tv = np.zeros((2, P.N), dtype=np.uint32) if export else None
        ksk = np.zeros((P.ksk_rows, P.n + 1), dtype=np.uint32) if (export and with_ksk) else None
        bsk = np.zeros((P.n, 2 * P.L, 2, P.N), dtype=np.float64) if export else None
        self._ck(self.lib.tfhe_ctx_generate_cloudkey(self.h, _ptr(s0), _ptr(s1), P.alpha_lv0, P.alpha_lv1, _seed(seed),
                                                     1 if with_ksk else 0, ctypes.cast(ctypes.byref(off), ctypes.c_void_p),
                                                     _ptr(bsk), _ptr(ksk), _ptr(tv)), "tfhe_ctx_generate_cloudkey")
        return (off.value, tv, ksk, bsk) if export else None

    def load_cloudkey_device(self, offset, d_bsk_fft, d_ksk, d_testvec, stream=0):
        self._ck(self.lib.tfhe_ctx_load_cloudkey_device(self.h, ctypes.c_uint32(offset), d_bsk_fft, d_ksk, d_testvec,
                                                        stream), "tfhe_ctx_load_cloudkey_device")

    # --- host-buffer hot path -----------------------------------------------------------------------
    def bootstrap_batch(self, ct_in, luts=None):
        P = self.P
        ct = _u32(ct_in, (-1, P.n + 1))
        out = np.empty_like(ct)
        l, nl = None, 0
        if luts is not None:
            l = _u32(luts, (-1, 2 * P.N))
            nl = len(l)
        self._ck(self.lib.tfhe_bootstrap_batch(self.h, len(ct), _ptr(ct), _ptr(l), nl, _ptr(out)), "tfhe_bootstrap_batch")
        return out

    def bootstrap_batch_indexed(self, ct_in, luts, lut_index):
        """LUT table [nluts][2][N] + one index per ciphertext (tfhe_bootstrap_batch_indexed)."""
        P = self.P
        ct = _u32(ct_in, (-1, P.n + 1))
        l = _u32(luts, (-1, 2 * P.N))
        idx = np.ascontiguousarray(lut_index, dtype=np.int32).ravel()
        assert len(idx) == len(ct)
        out = np.empty_like(ct)
        self._ck(self.lib.tfhe_bootstrap_batch_indexed(self.h, len(ct), _ptr(ct), _ptr(l), len(l), _ptr(idx), _ptr(out)),
                 "tfhe_bootstrap_batch_indexed")
        return out

    def bootstrap_multi_lut_batch(self, ct_in, packed_luts, log2_k):
        """2^log2_k functions per ciphertext from one blind rotation -> [count][k][n+1] (tfhe_bootstrap_multi_lut_batch)."""
        P = self.P
        ct = _u32(ct_in, (-1, P.n + 1))
        l = _u32(packed_luts, (-1, 2 * P.N))
        out = np.empty((len(ct), 1 << log2_k, P.n + 1), dtype=np.uint32)
        self._ck(self.lib.tfhe_bootstrap_multi_lut_batch(self.h, len(ct), _ptr(ct), _ptr(l), len(l), int(log2_k), _ptr(out)),
                 "tfhe_bootstrap_multi_lut_batch")
        return out

    def load_reencryption_key(self, key_encryptions, basebit, t):
        """proxyreenc.ProxyReencryptionKey.KeyEncryptions flattened [n*t*base][n+1]."""
        k = _u32(key_encryptions)
        assert k.size == self.P.n * t * (1 << basebit) * (self.P.n + 1)
        self._ck(self.lib.tfhe_ctx_load_reencryption_key(self.h, _ptr(k), int(basebit), int(t)), "tfhe_ctx_load_reencryption_key")

    def reencrypt_batch(self, ct_in):
        """proxyreenc.ReencryptTLWELv0 for a batch."""
        ct = _u32(ct_in, (-1, self.P.n + 1))
        out = np.empty_like(ct)
        self._ck(self.lib.tfhe_reencrypt_batch(self.h, len(ct), _ptr(ct), _ptr(out)), "tfhe_reencrypt_batch")
        return out

    def set_circuit_graph(self, enable=True):
        """CUDA-graph replay of repeated circuits in circuit_run (off by default; results identical)."""
        self._ck(self.lib.tfhe_ctx_set_circuit_graph(self.h, 1 if enable else 0), "tfhe_ctx_set_circuit_graph")

    @property
    def circuit_graph_replays(self):
        return int(self.lib.tfhe_ctx_circuit_graph_replays(self.h))

    def set_mux_mode(self, mode):
        """0 = the reference's three-bootstrap MUX (default), 1 = two blind rotations + one key switch (opt-in)."""
        self._ck(self.lib.tfhe_ctx_set_mux_mode(self.h, int(mode)), "tfhe_ctx_set_mux_mode")

    def gate_batch(self, ops, a, b=None, c=None):
        P = self.P
        a = _u32(a, (-1, P.n + 1))
        b = None if b is None else _u32(b, (-1, P.n + 1))
        c = None if c is None else _u32(c, (-1, P.n + 1))
        ov = _ops(ops, len(a))
        out = np.empty_like(a)
        self._ck(self.lib.tfhe_gate_batch(self.h, len(a), _ptr(ov), len(ov), _ptr(a), _ptr(b), _ptr(c), _ptr(out)),
                 "tfhe_gate_batch")
        return out

    def blind_rotate_batch(self, ct_in, luts=None):
        P = self.P
        ct = _u32(ct_in, (-1, P.n + 1))
        out = np.empty((len(ct), 2, P.N), dtype=np.uint32)
        l, nl = None, 0
        if luts is not None:
            l = _u32(luts, (-1, 2 * P.N))
            nl = len(l)
        self._ck(self.lib.tfhe_blind_rotate_batch(self.h, len(ct), _ptr(ct), _ptr(l), nl, _ptr(out)),
                 "tfhe_blind_rotate_batch")
        return out

    def cmux_batch(self, bsk_index, ct0, ct1):
        P = self.P
        ct1 = _u32(ct1, (-1, 2 * P.N))
        ct0 = None if ct0 is None else _u32(ct0, (-1, 2 * P.N))
        out = np.empty_like(ct1)
        self._ck(self.lib.tfhe_cmux_batch(self.h, len(ct1), int(bsk_index), _ptr(ct0), _ptr(ct1), _ptr(out)),
                 "tfhe_cmux_batch")
        return out.reshape(-1, 2, P.N)

    def sample_extract_batch(self, trlwe):
        P = self.P
        t = _u32(trlwe, (-1, 2 * P.N))
        out = np.empty((len(t), P.N + 1), dtype=np.uint32)
        self._ck(self.lib.tfhe_sample_extract_batch(self.h, len(t), _ptr(t), _ptr(out)), "tfhe_sample_extract_batch")
        return out

    def key_switch_batch(self, lwe1):
        P = self.P
        x = _u32(lwe1, (-1, P.N + 1))
        out = np.empty((len(x), P.n + 1), dtype=np.uint32)
        self._ck(self.lib.tfhe_key_switch_batch(self.h, len(x), _ptr(x), _ptr(out)), "tfhe_key_switch_batch")
        return out

    def to_fourier_batch(self, polys):
        """poly.Evaluator.ToFourierPoly for a batch: [count][N] u32 -> [count][N] f64 (reference layout)."""
        p = _u32(polys, (-1, self.P.N))
        out = np.empty(p.shape, dtype=np.float64)
        self._ck(self.lib.tfhe_to_fourier_batch(self.h, len(p), _ptr(p), _ptr(out)), "tfhe_to_fourier_batch")
        return out

    def to_poly_batch(self, fps):
        """poly.Evaluator.ToPoly for a batch: [count][N] f64 (reference layout) -> [count][N] u32."""
        f = np.ascontiguousarray(fps, dtype=np.float64).reshape(-1, self.P.N)
        out = np.empty(f.shape, dtype=np.uint32)
        self._ck(self.lib.tfhe_to_poly_batch(self.h, len(f), _ptr(f), _ptr(out)), "tfhe_to_poly_batch")
        return out

    def mul_poly_batch(self, p0, p1):
        """poly.Evaluator.MulPoly for a batch."""
        a, b = _u32(p0, (-1, self.P.N)), _u32(p1, (-1, self.P.N))
        out = np.empty_like(a)
        self._ck(self.lib.tfhe_mul_poly_batch(self.h, len(a), _ptr(a), _ptr(b), _ptr(out)), "tfhe_mul_poly_batch")
        return out

    def circuit_run(self, gates, n_inputs, inputs, output_wires):
        """gates: list of (op, in0, in1, in2, out); inputs [n_inputs][instances][n+1] -> [n_outputs][instances][n+1]."""
        P = self.P
        inputs = _u32(inputs).reshape(n_inputs, -1, P.n + 1)
        instances = inputs.shape[1]
        arr = (_native.GateDesc * max(len(gates), 1))()
        for k, (op, i0, i1, i2, o) in enumerate(gates):
            arr[k] = _native.GateDesc(OPCODES[op.upper()] if isinstance(op, str) else int(op), i0, i1, i2, o)
        ow = np.ascontiguousarray(output_wires, dtype=np.int32)
        out = np.empty((len(ow), instances, P.n + 1), dtype=np.uint32)
        self._ck(self.lib.tfhe_circuit_run(self.h, instances, n_inputs, len(gates), ctypes.cast(arr, ctypes.c_void_p),
                                           _ptr(inputs), len(ow), _ptr(ow), _ptr(out)), "tfhe_circuit_run")
        return out

    # --- device-buffer hot path (pointers are ints) ---------------------------------------------------
    def bootstrap_batch_device(self, count, d_ct_in, d_ct_out, d_luts=None, nluts=0, stream=0):
        self._ck(self.lib.tfhe_bootstrap_batch_device(self.h, count, d_ct_in, d_luts, nluts, d_ct_out, stream),
                 "tfhe_bootstrap_batch_device")

    def gate_batch_device(self, count, ops, d_a, d_b, d_c, d_out, stream=0):
        ov = _ops(ops, count)
        self._ck(self.lib.tfhe_gate_batch_device(self.h, count, _ptr(ov), len(ov), d_a, d_b, d_c, d_out, stream),
                 "tfhe_gate_batch_device")

    def set_blind_rotate_variant(self, variant):
        """'ldg' (automatic, default) | 'throughput' | 'lat' | 'latp'; the other names are round-1 experiments that exist
        only in a -DTFHE_EXPERIMENTAL=1 build.  Results are identical."""
        v = {"ldg": 0, "tma": 1, "tex": 2, "w16": 3, "tmem": 4, "tmex": 5, "tmex+tma": 6, "tms": 7, "mg": 8, "lat": 9, "throughput": 10, "lat2": 11, "latp": 12, "cl": 13}[variant] if isinstance(variant, str) else int(variant)
        self._ck(self.lib.tfhe_ctx_set_blind_rotate_variant(self.h, v), "tfhe_ctx_set_blind_rotate_variant")

    def set_blind_rotate_chunk_steps(self, steps):
        """CMUX steps per work item of the persistent throughput kernel (0 = automatic); results do not depend on it."""
        self._ck(self.lib.tfhe_ctx_set_blind_rotate_chunk_steps(self.h, int(steps)), "tfhe_ctx_set_blind_rotate_chunk_steps")

    def set_key_switch_variant(self, variant):
        """'auto' (default) | 'gather' | 'mma' | 'tile' — row gather out of L2, one tensor-core contraction (basebit = 2 sets) or
        shared-memory tiles of 256 ciphertexts (basebit >= 4 sets); results are identical."""
        v = {"auto": 0, "gather": 1, "mma": 2, "tile": 3}[variant] if isinstance(variant, str) else int(variant)
        self._ck(self.lib.tfhe_ctx_set_key_switch_variant(self.h, v), "tfhe_ctx_set_key_switch_variant")

    def set_timing(self, enable=True):
        self._ck(self.lib.tfhe_ctx_set_timing(self.h, 1 if enable else 0), "tfhe_ctx_set_timing")

    def collect_timing(self):
        """{'blind_rotate_ms', 'blind_rotate_launches', 'key_switch_ms', 'key_switch_launches'} since last collect."""
        out = (ctypes.c_double * 4)()
        self._ck(self.lib.tfhe_ctx_collect_timing(self.h, ctypes.byref(out)), "tfhe_ctx_collect_timing")
        return {"blind_rotate_ms": out[0], "blind_rotate_launches": int(out[1]), "key_switch_ms": out[2],
                "key_switch_launches": int(out[3])}

    @property
    def kernel_launches(self):
        return int(self.lib.tfhe_ctx_kernel_launches(self.h))

    @property
    def algorithmic_bytes_per_bootstrap(self):
        return int(self.lib.tfhe_ctx_algorithmic_bytes_per_bootstrap(self.h))
