"""Parameter sets — mirrors params/params.go:83-391 of the reference (Uint6-8 are not offered: the
reference itself marks them broken, params/UINT_STATUS.md:14-31).

Unlike the reference's mutable package global (params.CurrentSecurityLevel, params.go:47) a ParamSet
is an immutable value captured by every key and context."""
from dataclasses import dataclass


@dataclass(frozen=True)
class ParamSet:
    name: str
    n: int            # TLWELv0.N
    alpha_lv0: float  # TLWELv0.ALPHA  (KSKAlpha, params.go:629)
    N: int            # TRGSWLv1.N
    alpha_lv1: float  # TLWELv1.ALPHA  (BSKAlpha, params.go:634)
    NBIT: int
    BGBIT: int
    L: int
    BASEBIT: int
    IKS_T: int

    @property
    def BG(self):
        return 1 << self.BGBIT

    @property
    def base(self):
        return 1 << self.BASEBIT

    @property
    def ksk_rows(self):
        return self.N * self.IKS_T * self.base

    @property
    def algorithmic_bytes_per_bootstrap(self):
        """BK rows + expected KSK rows + 2 ct in / 1 out."""
        return (self.n * 2 * self.L * 2 * self.N * 8 + self.N * self.IKS_T * (self.base - 1) // self.base * (self.n + 1) * 4
                + 3 * (self.n + 1) * 4)

    @property
    def flops_per_bootstrap(self):
        """SURVEY.md section 8(d) secondary roof: radix-2 flop count of the reference algorithm."""
        import math
        M = self.N // 2
        fft = M // 2 * int(math.log2(M)) * 10
        step = 2 * self.L * fft + 2 * (fft + self.N) + 4 * self.L * M * 8 + 2 * self.N * 5
        return self.n * step


_SETS = [
    ParamSet("80", 550, 5.0e-5, 1024, 3.73e-8, 10, 6, 3, 2, 7),
    ParamSet("110", 630, 3.0517578125e-05, 1024, 2.980232238769531e-8, 10, 6, 3, 2, 8),
    ParamSet("128", 700, 2.0e-5, 1024, 2.0e-8, 10, 6, 3, 2, 9),
    ParamSet("uint1", 700, 2.0e-05, 1024, 2.0e-08, 10, 10, 2, 2, 8),
    ParamSet("uint2", 687, 0.00002120846893069971872305794214, 512, 0.00000000000231841227527049948463, 9, 18, 1, 4, 3),
    ParamSet("uint3", 820, 0.00000251676160959795544987084234, 1024, 0.00000000000000022204460492503131, 10, 23, 1, 6, 2),
    ParamSet("uint4", 820, 0.00000251676160959795544987084234, 2048, 0.00000000000000022204460492503131, 11, 22, 1, 5, 3),
    ParamSet("uint5", 1071, 7.088226765410429399593757e-08, 2048, 2.2204460492503131e-17, 11, 22, 1, 6, 3),
]
SETS = {p.name: p for p in _SETS}
Security80Bit, Security110Bit, Security128Bit = "80", "110", "128"
SecurityUint1, SecurityUint2, SecurityUint3, SecurityUint4, SecurityUint5 = "uint1", "uint2", "uint3", "uint4", "uint5"

# Default level used when a caller does not pass one (the reference's default, params.go:47).
CurrentSecurityLevel = Security128Bit


def get(name=None):
    return SETS[str(name if name is not None else CurrentSecurityLevel)]
