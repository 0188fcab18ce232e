"""lut — mirrors lut/generator.go and lut/lut.go (host side, tiny)."""
import ctypes

import numpy as np

from . import _native
from .key import _tp


class LookUpTable:
    """lut.LookUpTable (lut/lut.go:14-17): Poly is a TRLWE [2][N] with A = 0."""

    def __init__(self, poly):
        self.Poly = poly


class Generator:
    """lut.Generator (lut/generator.go:15-39)."""

    def __init__(self, messageModulus, P):
        self.MessageModulus, self.P = int(messageModulus), P

    def GenLookUpTable(self, f):
        """lut/generator.go:49-100."""
        fv = np.array([int(f(x)) for x in range(self.MessageModulus)], dtype=np.int32)
        out = np.zeros((2, self.P.N), dtype=np.uint32)
        _native.client().tfhe_client_gen_lut(ctypes.byref(_tp(self.P)), self.MessageModulus, fv.ctypes.data,
                                             out.ctypes.data)
        return LookUpTable(out)


def NewGenerator(messageModulus, P):
    return Generator(messageModulus, P)


def PackLookUpTables(luts):
    """Test vector of a many-LUT bootstrap (tfhe_bootstrap_multi_lut_batch): k = len(luts) (a power of two) LookUpTables of
    the same message modulus interleaved coefficient-wise, packed[j] = luts[j mod k][j].  Returns a TRLWE [2][N]."""
    k = len(luts)
    assert k >= 2 and k & (k - 1) == 0, "the number of functions must be a power of two"
    polys = np.stack([np.asarray(l.Poly if hasattr(l, "Poly") else l, dtype=np.uint32).reshape(2, -1) for l in luts])
    N = polys.shape[2]
    j = np.arange(N)
    return np.ascontiguousarray(polys[j % k, :, j].T)
