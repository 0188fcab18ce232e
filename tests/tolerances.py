"""Stated tolerances for the parameter sets whose f64 external products are NOT exact (L = 1 with a 2^18..2^23 gadget base:
the reference's own sums exceed 2^53, so its rounded result depends on summation order and GPU == oracle == Go only up to
a tolerance).  Values are PHASE differences in torus LSB (2^-32 of the torus), per set, about 4x the worst case observed
GPU-vs-oracle on a B200 (gpurun_out/r02_uint_tolerances.txt, round 2: identity / complement / modulo LUTs over every message):

    set     observed after blind rotate   observed after key switch     decode half-slot (2^31 / msgMod)
    uint1   0                              0                             (exact: L = 2, Bg = 2^10 keeps the sums below 2^53)
    uint2   4 374 840  (2^22.1)            12 205 591 (2^23.5)           536 870 912
    uint3     220 368  (2^17.7)            19 516 480 (2^24.2)           268 435 456
    uint4     869 238  (2^19.7)             4 686 995 (2^22.2)           134 217 728
    uint5   1 072 042  (2^20.0)             1 101 431 (2^20.1)            67 108 864

After the key switch the two sides carry DIFFERENT masks (one LSB of difference before a digit boundary swaps in another key
row), so their phases differ by two independent key-switch rounding noises: that column is a noise bound, not an
arithmetic error.  Every tolerance is far inside the decode margin of its set, and decoded messages must be equal anyway."""

UINT_PHASE_TOL = {           # (after blind rotate + sample extract, after key switch)
    "uint1": (0, 0),
    "uint2": (1 << 24, 3 << 24),
    "uint3": (1 << 20, 5 << 24),
    "uint4": (1 << 22, 5 << 22),
    "uint5": (1 << 22, 1 << 22),
}
