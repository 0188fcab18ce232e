"""key.SecretKey — mirrors key/key.go:10-45 (client side; runs on the host)."""
import ctypes

import numpy as np

from . import _native, params as _params


def _tp(P):
    return _native.TfheParams(P.n, P.N, P.L, P.BGBIT, P.BASEBIT, P.IKS_T)


class SecretKey:
    def __init__(self, P, KeyLv0, KeyLv1):
        self.P, self.KeyLv0, self.KeyLv1 = P, KeyLv0, KeyLv1


def NewSecretKey(P=None, seed=0):
    """key.NewSecretKey (key/key.go:16-45).  The reference draws from unseeded math/rand; here the seed is explicit."""
    P = P or _params.get()
    s0 = np.zeros(P.n, dtype=np.uint32)
    s1 = np.zeros(P.N, dtype=np.uint32)
    _native.client().tfhe_client_secret_key(ctypes.byref(_tp(P)), seed, s0.ctypes.data, s1.ctypes.data)
    return SecretKey(P, s0, s1)
