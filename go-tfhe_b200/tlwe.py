"""tlwe — LWE ciphertext helpers, mirrors tlwe/tlwe.go and tlwe/programmable_encrypt.go (client side).
A ciphertext is a uint32 array of n+1 words (P[0..n-1] = mask, P[n] = b); batches are [count][n+1]."""
import ctypes

import numpy as np

from . import _native
from .key import _tp, _seed


def EncryptBool(bits, sk, seed=None, alpha=None):
    """tlwe.EncryptBool (tlwe/tlwe.go:54-62) for every element of `bits`."""
    P = sk.P
    bits = np.ascontiguousarray(bits, dtype=np.uint8).ravel()
    out = np.zeros((len(bits), P.n + 1), dtype=np.uint32)
    _native.client().tfhe_client_encrypt_bool(ctypes.byref(_tp(P)), alpha if alpha is not None else P.alpha_lv0,
                                              sk.KeyLv0.ctypes.data, _seed(seed), len(bits), bits.ctypes.data, out.ctypes.data)
    return out


def DecryptBool(ct, sk):
    """tlwe.DecryptBool (tlwe/tlwe.go:65-74)."""
    P = sk.P
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, P.n + 1)
    bits = np.zeros(len(ct), dtype=np.uint8)
    _native.client().tfhe_client_decrypt_bool(ctypes.byref(_tp(P)), sk.KeyLv0.ctypes.data, len(ct), ct.ctypes.data,
                                              bits.ctypes.data)
    return bits


def EncryptLWEMessage(msgs, messageModulus, sk, seed=None, alpha=None):
    """tlwe.EncryptLWEMessage (tlwe/programmable_encrypt.go:12-27)."""
    P = sk.P
    msgs = np.ascontiguousarray(msgs, dtype=np.int32).ravel()
    out = np.zeros((len(msgs), P.n + 1), dtype=np.uint32)
    _native.client().tfhe_client_encrypt_message(ctypes.byref(_tp(P)), alpha if alpha is not None else P.alpha_lv0,
                                                 sk.KeyLv0.ctypes.data, _seed(seed), len(msgs), msgs.ctypes.data,
                                                 int(messageModulus), out.ctypes.data)
    return out


def DecryptLWEMessage(ct, messageModulus, sk):
    """tlwe.DecryptLWEMessage (tlwe/programmable_encrypt.go:33-54)."""
    P = sk.P
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, P.n + 1)
    out = np.zeros(len(ct), dtype=np.int32)
    _native.client().tfhe_client_decrypt_message(ctypes.byref(_tp(P)), sk.KeyLv0.ctypes.data, len(ct), ct.ctypes.data,
                                                 int(messageModulus), out.ctypes.data)
    return out


def Neg(ct):
    """TLWELv0.Neg (tlwe/tlwe.go:103-109)."""
    return (np.uint32(0) - np.asarray(ct, dtype=np.uint32)).astype(np.uint32)
