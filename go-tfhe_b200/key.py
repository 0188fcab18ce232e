"""key.SecretKey — mirrors key/key.go:10-45 (client side; runs on the host)."""
import ctypes

import numpy as np

from . import _native, params as _params


def _tp(P):
    return _native.TfheParams(P.n, P.N, P.L, P.BGBIT, P.BASEBIT, P.IKS_T)


class SecretKey:
    def __init__(self, P, KeyLv0, KeyLv1):
        self.P, self.KeyLv0, self.KeyLv1 = P, KeyLv0, KeyLv1


def _seed(seed):
    """None -> 0 = 'draw a 256-bit key from the OS' for the C side; integers are reproducible test seeds (0 is mapped away)."""
    if seed is None:
        return 0
    seed = int(seed) & (2**64 - 1)
    return seed if seed != 0 else 0x9E3779B97F4A7C15


def NewSecretKey(P=None, seed=None):
    """key.NewSecretKey (key/key.go:16-45).  seed=None (default): fresh key bits from ChaCha20 under an OS-entropy key, like
    the reference's self-seeding RNG but cryptographically strong; an integer seed (non-zero) gives a reproducible key for tests."""
    P = P or _params.get()
    s0 = np.zeros(P.n, dtype=np.uint32)
    s1 = np.zeros(P.N, dtype=np.uint32)
    _native.client().tfhe_client_secret_key(ctypes.byref(_tp(P)), _seed(seed), s0.ctypes.data, s1.ctypes.data)
    return SecretKey(P, s0, s1)
