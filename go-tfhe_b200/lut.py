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
