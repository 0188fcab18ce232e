"""evaluator.Evaluator — mirrors evaluator/evaluator.go and evaluator/programmable_bootstrap.go.
Same method names and argument meaning as the reference; every method also accepts a batch
([count][n+1]) where the reference takes one ciphertext, and is applied element-wise on the GPU.
The reference's bsk / ksk / decompositionOffset arguments are carried by the CloudKey."""
import numpy as np

from . import lut as _lut


class Evaluator:
    def __init__(self, cloudKey, device=0):
        """evaluator.NewEvaluator (evaluator/evaluator.go:27-35): scratch lives on the GPU context."""
        self.ck = cloudKey
        self.ctx = cloudKey.engine(device)
        self.P = cloudKey.P

    def _shape(self, ct):
        ct = np.asarray(ct, dtype=np.uint32)
        return ct, ct.ndim == 1

    def ExternalProduct(self, bskIndex, ctIn):
        """ExternalProductAssign (evaluator.go:50-81) with BootstrappingKey[bskIndex]; ctIn TRLWE [..][2][N]."""
        x = np.asarray(ctIn, dtype=np.uint32)
        out = self.ctx.cmux_batch(bskIndex, None, x)
        return out[0] if x.ndim == 2 else out

    def CMux(self, bskIndex, ct0, ct1):
        """CMuxAssign (evaluator.go:85-106): ct0 + BootstrappingKey[bskIndex] (x) (ct1 - ct0)."""
        x = np.asarray(ct1, dtype=np.uint32)
        out = self.ctx.cmux_batch(bskIndex, ct0, x)
        return out[0] if x.ndim == 2 else out

    def BlindRotate(self, ctIn, testvec=None):
        """BlindRotateAssign (evaluator.go:110-135).  testvec None => CloudKey.BlindRotateTestvec."""
        ct, single = self._shape(ctIn)
        out = self.ctx.blind_rotate_batch(ct, testvec)
        return out[0] if single else out

    def Bootstrap(self, ctIn, testvec=None):
        """Bootstrap / BootstrapAssign (evaluator.go:139-157)."""
        ct, single = self._shape(ctIn)
        out = self.ctx.bootstrap_batch(ct, testvec)
        return out[0] if single else out

    def BootstrapLUT(self, ctIn, lut):
        """BootstrapLUT / BootstrapLUTAssign (programmable_bootstrap.go:54-115).  lut: LookUpTable, or a
        list of LookUpTables (one per ciphertext)."""
        polys = lut.Poly if isinstance(lut, _lut.LookUpTable) else np.stack([l.Poly for l in lut])
        return self.Bootstrap(ctIn, polys)

    def BootstrapFunc(self, ctIn, f, messageModulus):
        """BootstrapFunc (programmable_bootstrap.go:16-29)."""
        return self.BootstrapLUT(ctIn, _lut.NewGenerator(messageModulus, self.P).GenLookUpTable(f))

    def SampleExtractIndex0(self, trlwe):
        """trlwe.SampleExtractIndexAssign(k=0) (trlwe/trlwe_ops.go:10-21)."""
        x = np.asarray(trlwe, dtype=np.uint32)
        out = self.ctx.sample_extract_batch(x)
        return out[0] if x.ndim == 2 else out

    def IdentityKeySwitching(self, lwe1):
        """trgsw.IdentityKeySwitchingAssign (trgsw/keyswitch.go:10-37)."""
        x = np.asarray(lwe1, dtype=np.uint32)
        out = self.ctx.key_switch_batch(x)
        return out[0] if x.ndim == 1 else out


def NewEvaluator(cloudKey, device=0):
    return Evaluator(cloudKey, device)
