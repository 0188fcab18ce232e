"""gates — mirrors gates/gates.go.  gates.X(a, b, ck) keeps the reference signature; a and b may be single
ciphertexts ([n+1]) or batches ([count][n+1]).  gates.BatchX(inputs, ck) takes the reference's list of pairs.
Every element of a batch equals the single-gate path (the reference's intended semantics; its BatchXNOR
bias sign and result aliasing defects are not reproduced)."""
import numpy as np

from . import tlwe as _tlwe

Ciphertext = np.ndarray  # gates.Ciphertext = tlwe.TLWELv0 (gates/gates.go:16)


def _gate(op, a, b, ck, c=None, device=0):
    a = np.asarray(a, dtype=np.uint32)
    single = a.ndim == 1
    out = ck.engine(device).gate_batch(op, a, b, c)
    return out[0] if single else out


def NAND(a, b, ck): return _gate("NAND", a, b, ck)      # gates.go:26-31
def OR(a, b, ck): return _gate("OR", a, b, ck)          # gates.go:34-37
def AND(a, b, ck): return _gate("AND", a, b, ck)        # gates.go:40-43
def XOR(a, b, ck): return _gate("XOR", a, b, ck)        # gates.go:46-49
def XNOR(a, b, ck): return _gate("XNOR", a, b, ck)      # gates.go:52-58
def NOR(a, b, ck): return _gate("NOR", a, b, ck)        # gates.go:72-76
def ANDNY(a, b, ck): return _gate("ANDNY", a, b, ck)    # gates.go:79-83
def ANDYN(a, b, ck): return _gate("ANDYN", a, b, ck)    # gates.go:86-90
def ORNY(a, b, ck): return _gate("ORNY", a, b, ck)      # gates.go:93-97
def ORYN(a, b, ck): return _gate("ORYN", a, b, ck)      # gates.go:100-104
def MUX(a, b, c, ck): return _gate("MUX", a, b, ck, c)  # gates.go:107-114: a ? b : c, three bootstraps


def NOT(a):
    """gates.NOT (gates.go:117-119): no bootstrap, pure host arithmetic exactly as in the reference."""
    return _tlwe.Neg(a)


def Copy(a):
    """gates.Copy (gates.go:122-126)."""
    return np.array(a, dtype=np.uint32, copy=True)


def Constant(value, P):
    """gates.Constant (gates.go:61-69): trivial ciphertext (0,...,0, mu), mu = 1/8 or 1 - 1/8 (sic, uint32)."""
    out = np.zeros(P.n + 1, dtype=np.uint32)
    out[P.n] = 0x20000000 if value else (1 - 0x20000000) & 0xFFFFFFFF
    return out


def _batch(op, inputs, ck):
    a = np.stack([np.asarray(p[0], dtype=np.uint32) for p in inputs])
    b = np.stack([np.asarray(p[1], dtype=np.uint32) for p in inputs])
    return list(_gate(op, a, b, ck))


def BatchNAND(inputs, ck): return _batch("NAND", inputs, ck)  # gates.go:156-182
def BatchAND(inputs, ck): return _batch("AND", inputs, ck)    # gates.go:185-208
def BatchOR(inputs, ck): return _batch("OR", inputs, ck)      # gates.go:211-234
def BatchXOR(inputs, ck): return _batch("XOR", inputs, ck)    # gates.go:237-260
def BatchNOR(inputs, ck): return _batch("NOR", inputs, ck)    # gates.go:263-286
def BatchXNOR(inputs, ck): return _batch("XNOR", inputs, ck)  # gates.go:289-312 (with the single-gate bias)


def BatchMixed(ops, a, b, c, ck, device=0):
    """Additive API: one opcode per gate, MUX included."""
    return ck.engine(device).gate_batch(ops, a, b, c)
