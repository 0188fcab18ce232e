"""cloudkey.CloudKey — mirrors cloudkey/cloudkey.go:16-145.  Fields keep the reference's names and
meaning but are flat numpy arrays (the layouts of include/tfhe_b200.h), which is exactly what the cgo shim
produces from the Go structs before crossing the C ABI."""
import ctypes
import os

import numpy as np

from . import _native, engine
from .key import _seed, _tp


class CloudKey:
    def __init__(self, P, DecompositionOffset, BlindRotateTestvec, KeySwitchingKey, BootstrappingKey):
        self.P = P
        self.DecompositionOffset = int(DecompositionOffset)
        self.BlindRotateTestvec = BlindRotateTestvec      # [2][N] u32
        self.KeySwitchingKey = KeySwitchingKey            # [N*t*base][n+1] u32 or None
        self.BootstrappingKey = BootstrappingKey          # [n][2L][2][N] f64, reference FourierPoly layout
        self._ctx = {}

    def engine(self, device=0):
        """The GPU context holding this key (created and uploaded on first use)."""
        if device not in self._ctx:
            ctx = engine.Context(self.P, device)
            ctx.load_cloudkey(self.DecompositionOffset, self.BootstrappingKey, self.KeySwitchingKey,
                              self.BlindRotateTestvec)
            self._ctx[device] = ctx
        return self._ctx[device]

    def close(self):
        for c in self._ctx.values():
            c.close()
        self._ctx = {}


def NewCloudKey(secretKey, seed=None, threads=None, with_ksk=True):
    """cloudkey.NewCloudKey (cloudkey/cloudkey.go:24-31); with_ksk=False ~ NewCloudKeyNoKSK (:34-57)."""
    P = secretKey.P
    off = ctypes.c_uint32(0)
    tv = np.zeros((2, P.N), dtype=np.uint32)
    ksk = np.zeros((P.ksk_rows, P.n + 1), dtype=np.uint32) if with_ksk else None
    bsk = np.zeros((P.n, 2 * P.L, 2, P.N), dtype=np.float64)
    _native.client().tfhe_client_cloud_key(ctypes.byref(_tp(P)), P.alpha_lv0, P.alpha_lv1, secretKey.KeyLv0.ctypes.data,
                                           secretKey.KeyLv1.ctypes.data, _seed(seed), threads or (os.cpu_count() or 1),
                                           ctypes.byref(off), tv.ctypes.data, ksk.ctypes.data if with_ksk else None,
                                           bsk.ctypes.data)
    return CloudKey(P, off.value, tv, ksk, bsk)


def NewCloudKeyOnDevice(secretKey, seed=None, device=0, with_ksk=True, export=True):
    """cloudkey.NewCloudKey (cloudkey/cloudkey.go:24-31) evaluated on the GPU (tfhe_ctx_generate_cloudkey): the
    bootstrapping and key-switching keys are produced in device memory and left loaded in the engine; with export=True
    the returned CloudKey also carries the reference-layout fields (what a Go caller would store), otherwise only the
    engine handle."""
    P = secretKey.P
    ctx = engine.Context(P, device)
    res = ctx.generate_cloudkey(secretKey.KeyLv0, secretKey.KeyLv1, seed, with_ksk, export)
    if export:
        off, tv, ksk, bsk = res
        ck = CloudKey(P, off, tv, ksk, bsk)
    else:
        ck = CloudKey(P, 0, None, None, None)
    ck._ctx[device] = ctx
    return ck
