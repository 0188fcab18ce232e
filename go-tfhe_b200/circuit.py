"""Levelised circuits over the gate API — the caller directly above the hot path (README.md:78-114 of the
reference builds its adders gate by gate through gates.*; here the same gate list runs level by level with all
intermediate ciphertexts resident on the GPU: tfhe_circuit_run)."""
import numpy as np

from . import gates as _gates


class Circuit:
    def __init__(self, n_inputs):
        self.n_inputs = n_inputs
        self.n_wires = n_inputs
        self.gates = []          # (op, in0, in1, in2, out)
        self.constants = {}      # input wire -> bool (trivial ciphertexts appended by run())
        self.out_wires = []

    def constant(self, value):
        """gates.Constant (gates/gates.go:61-69) as an extra input wire.  Call before adding gates."""
        assert not self.gates, "declare constants before gates"
        w = self.n_wires
        self.n_wires += 1
        self.constants[w] = bool(value)
        return w

    def gate(self, op, a, b=None, c=None):
        out = self.n_wires
        self.n_wires += 1
        self.gates.append((op.upper(), a, a if b is None else b, 0 if c is None else c, out))
        return out

    def outputs(self, wires):
        self.out_wires = list(wires)
        return self

    @property
    def n_levels(self):
        depth = {w: 0 for w in range(self.n_inputs + len(self.constants))}
        for op, a, b, c, o in self.gates:
            if op in ("NOT", "COPY"):
                depth[o] = depth[a]
            elif op == "MUX":
                depth[o] = 2 + max(depth[a], depth[b], depth[c])
            else:
                depth[o] = 1 + max(depth[a], depth[b])
        return max(depth.values()) if depth else 0

    @property
    def n_bootstraps(self):
        return sum(3 if g[0] == "MUX" else 0 if g[0] in ("NOT", "COPY") else 1 for g in self.gates)

    def run(self, ck, inputs, device=0):
        """inputs: [n_inputs][instances][n+1] ciphertexts -> [n_outputs][instances][n+1]."""
        P = ck.P
        inputs = np.asarray(inputs, dtype=np.uint32).reshape(self.n_inputs, -1, P.n + 1)
        inst = inputs.shape[1]
        if self.constants:
            extra = np.stack([np.broadcast_to(_gates.Constant(v, P), (inst, P.n + 1)) for _, v in sorted(self.constants.items())])
            inputs = np.concatenate([inputs, extra])
        return ck.engine(device).circuit_run(self.gates, inputs.shape[0], inputs, self.out_wires)


def full_adder(c, a, b, cin):
    """FullAdder of the reference's README.md:78-87: XOR, AND, AND, XOR, OR."""
    a_xor_b = c.gate("XOR", a, b)
    a_and_b = c.gate("AND", a, b)
    a_xor_b_and_c = c.gate("AND", a_xor_b, cin)
    s = c.gate("XOR", a_xor_b, cin)
    carry = c.gate("OR", a_and_b, a_xor_b_and_c)
    return s, carry


def ripple_carry_adder(bits):
    """README.md:89-114: inputs are wires 0..bits-1 (x, LSB first) and bits..2*bits-1 (y); carry-in Constant(false);
    outputs are the `bits` sum wires (mod 2^bits).  5 bootstraps per bit."""
    c = Circuit(2 * bits)
    carry = c.constant(False)
    sums = []
    for i in range(bits):
        s, carry = full_adder(c, i, bits + i, carry)
        sums.append(s)
    return c.outputs(sums)
