// The throughput blind-rotation kernel's shipped instantiations, as their own translation unit.
//
// Reason: register allocation.  The kernel sits at 252 registers with two warps per scheduler, and how ptxas spends those
// registers decides how much of the FP64 / shared-memory latency the two warps can cover.  `-Xptxas
// --register-usage-level=7` (default 5) gives the same instruction count with a different allocation and schedule:
// 40.46 -> 40.16 ms per 4096 gates at 128-bit (+0.7 %), 12.54 -> 11.99 ms per 2048 Uint3 bootstraps (+4.6 %), neutral at
// N = 2048.  The same option slows the latency kernels (single gate 2.49 -> 2.53 ms) and the tiled key switch (1.93 -> 2.03
// ms), so it is applied to this file only (go-tfhe_b200/build.py).  Results are bit-identical: same arithmetic.
#include "blind_rotate.cuh"

namespace tfhe {
// address of the instantiation for a parameter-set shape (nullptr: none); the engine launches through this pointer
void (*blind_rotate_throughput_kernel(int logN, int L, int bgbit))(const BrArgs) {
#define X(LOGN, L_, BG, SMALL, MINB) \
  if (logN == LOGN && L == L_ && bgbit == BG) return blind_rotate_kernel<LOGN, L_, BG, SMALL, MINB, false>;
  TFHE_BR_THROUGHPUT_INSTANCES(X)
#undef X
  return nullptr;
}
}  // namespace tfhe
