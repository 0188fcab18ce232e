// tmem_bench.cu — can TMEM (tcgen05.ld/st) serve as per-thread spill space for FP64 accumulators?
// Measures correctness of a register -> TMEM -> register round trip and the sustained ld / st throughput per SM
// with 4 warps per CTA (one TMEM lane quadrant each) and 1..4 CTAs per SM (128 columns each).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define COLS 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> struct Regs { uint32_t r[X]; };

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
               "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: correctness; 1: ld throughput; 2: st throughput; 3: ld+st (read-modify-write) with FP64 FMA in between
__global__ void __launch_bounds__(128) tmem_kernel(int mode, int iters, unsigned long long* cycles, uint32_t* errors) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_base)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);  // this warp's lane quadrant
  uint32_t r[16];
  if (mode == 0) {
    uint32_t bad = 0;
    for (int c = 0; c < COLS; c += 16) {
      for (int k = 0; k < 16; k++) r[k] = (blockIdx.x * 131u + threadIdx.x) * 1000u + c + k;
      tmem_st16(base + c, r);
    }
    tmem_wait_st();
    for (int c = 0; c < COLS; c += 16) {
      tmem_ld16(base + c, r);
      tmem_wait_ld();
      for (int k = 0; k < 16; k++) bad += (r[k] != (blockIdx.x * 131u + threadIdx.x) * 1000u + c + k);
    }
    if (bad) atomicAdd(errors, bad);
  } else {
    for (int k = 0; k < 16; k++) r[k] = lane + k;
    for (int c = 0; c < COLS; c += 16) tmem_st16(base + c, r);
    tmem_wait_st();
    __syncthreads();
    unsigned long long t0 = clock64();
    uint32_t acc = 0;
    double d0 = lane, d1 = 1.0000001;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int c = 0; c < COLS; c += 16) {
        if (mode == 1 || mode == 3) { tmem_ld16(base + c, r); tmem_wait_ld(); }
        if (mode == 3) {
#pragma unroll
          for (int k = 0; k < 16; k += 2) {
            double v = __hiloint2double(r[k + 1], r[k]);
            v = fma(v, d1, d0);
            r[k] = __double2loint(v); r[k + 1] = __double2hiint(v);
          }
        }
        if (mode == 2 || mode == 3) tmem_st16(base + c, r);
        if (mode == 1) acc += r[0];
      }
      if (mode == 2 || mode == 3) tmem_wait_st();
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0xdeadbeef) errors[1] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "n"(COLS) : "memory");
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  unsigned long long* d_cyc; uint32_t* d_err;
  cudaMalloc(&d_cyc, sizeof(unsigned long long) * sms * 8); cudaMalloc(&d_err, 8); cudaMemset(d_err, 0, 8);
  tmem_kernel<<<sms * 4, 128>>>(0, 1, d_cyc, d_err);
  uint32_t err[2]; cudaMemcpy(err, d_err, 8, cudaMemcpyDeviceToHost);
  printf("round trip: %s (errors=%u) [%s]\n", err[0] ? "FAIL" : "ok", err[0], cudaGetErrorString(cudaGetLastError()));
  const char* names[] = {"", "ld only", "st only", "ld+fma+st"};
  for (int mode = 1; mode <= 3; mode++)
    for (int per_sm = 1; per_sm <= 4; per_sm *= 2) {
      const int iters = 2000, grid = sms * per_sm;
      tmem_kernel<<<grid, 128>>>(mode, iters, d_cyc, d_err);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long c[148 * 4]; cudaMemcpy(c, d_cyc, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
      double mean = 0; for (int i = 0; i < grid; i++) mean += c[i]; mean /= grid;
      const double bytes_per_cta = (double)iters * 128 /*threads*/ * COLS * 4 * (mode == 3 ? 2 : 1);
      printf("%-10s %d CTA/SM (%2d warps): %.0f cycles, %.1f B/clk/SM  [%s]\n", names[mode], per_sm, per_sm * 4, mean,
             bytes_per_cta * per_sm / mean, cudaGetErrorString(e));
    }
  return 0;
}
