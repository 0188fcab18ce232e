"""poly.Evaluator — mirrors poly/poly_evaluator.go, poly/fourier_transform.go, poly/poly_mul.go for the transforms the
bootstrap path is made of.  FourierPoly values are float64 arrays of length N in the reference's own layout (groups of
4 real + 4 imaginary parts, poly/poly.go:54-62); everything runs on the GPU context."""
import numpy as np


class Evaluator:
    def __init__(self, ctx):
        """poly.NewEvaluator(N) (poly/poly_evaluator.go:76): twiddles and scratch live in the GPU context."""
        self.ctx = ctx
        self.N = ctx.P.N

    def _b(self, a, dtype):
        a = np.asarray(a, dtype=dtype)
        return a, a.ndim == 1

    def ToFourierPoly(self, p):
        """poly/fourier_transform.go:11-21."""
        p, single = self._b(p, np.uint32)
        out = self.ctx.to_fourier_batch(p)
        return out[0] if single else out

    def ToPoly(self, fp):
        """poly/fourier_transform.go:24-36."""
        fp, single = self._b(fp, np.float64)
        out = self.ctx.to_poly_batch(fp)
        return out[0] if single else out

    def MulPoly(self, p0, p1):
        """poly/poly_mul.go:4-22."""
        p0, single = self._b(p0, np.uint32)
        out = self.ctx.mul_poly_batch(p0, p1)
        return out[0] if single else out


def NewEvaluator(ctx):
    return Evaluator(ctx)
