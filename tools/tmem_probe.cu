// tmem_probe.cu — prints which (TMEM lane, column) each (thread, register) of a warp receives for the
// tcgen05.ld shapes, after the data was written with the 32x32b shape (thread t -> lane t).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t s_base;
  const int lane = threadIdx.x;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_base)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base;
  uint32_t r[16];
  for (int c0 = 0; c0 < 32; c0 += 16) {
    for (int k = 0; k < 16; k++) r[k] = (lane << 8) | (c0 + k);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(base + c0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  uint32_t a[8];
  // shape 0: 16x64b.x2 (2 regs), lane offset 0
  asm volatile("tcgen05.ld.sync.aligned.16x64b.x2.b32 {%0,%1}, [%2];" : "=r"(a[0]), "=r"(a[1]) : "r"(base) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 2; k++) out[(0 * 32 + lane) * 8 + k] = a[k];
  // shape 1: 16x128b.x2 (4 regs)
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 4; k++) out[(1 * 32 + lane) * 8 + k] = a[k];
  // shape 2: 16x256b.x2 (8 regs)
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]) : "r"(base) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 8; k++) out[(2 * 32 + lane) * 8 + k] = a[k];
  // shape 3: 16x256b.x1 at lane offset 16
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base + (16u << 16)) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 4; k++) out[(3 * 32 + lane) * 8 + k] = a[k];
  // shape 4: 16x32bx2.x2 with half split offset 8 columns
  asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x2.b32 {%0,%1}, [%2], 8;" : "=r"(a[0]), "=r"(a[1]) : "r"(base) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 2; k++) out[(4 * 32 + lane) * 8 + k] = a[k];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(s_base) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 5 * 32 * 8 * 4); cudaMemset(d, 0xff, 5 * 32 * 8 * 4);
  probe<<<1, 32>>>(d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  uint32_t h[5 * 32 * 8]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  const char* names[] = {"16x64b.x2", "16x128b.x2", "16x256b.x2", "16x256b.x1 @lane16", "16x32bx2.x2 split 8"};
  const int nreg[] = {2, 4, 8, 4, 2};
  for (int s = 0; s < 5; s++) {
    printf("== %s: thread: (lane,col) per register\n", names[s]);
    for (int t = 0; t < 32; t++) {
      printf("t%02d:", t);
      for (int k = 0; k < nreg[s]; k++) printf(" (%2u,%2u)", h[(s * 32 + t) * 8 + k] >> 8, h[(s * 32 + t) * 8 + k] & 0xff);
      printf("\n");
    }
  }
  return 0;
}
