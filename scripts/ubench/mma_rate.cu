// Microbenchmark: issue rate of legacy mma.sync.m16n8k16 (bf16) and ldmatrix.x4 on sm_100a, per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k(int iters, long long* out, int mode) {
  __shared__ __align__(16) uint8_t sm[16 * 2064 + 4096];
  for (int i = threadIdx.x; i < (int)sizeof(sm) / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x3c003c00u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm) + (lane & 15) * 2064 + (warp * 64 + (lane >> 4) * 8) * 2;
  float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode & 1) {
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr + (i & 3) * 32));
    }
    if (mode & 2) {
#define MMA(c) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
      if (mode & 4) { MMA(c0) MMA(c0) MMA(c0) MMA(c0) }   // dependent chain
      else { MMA(c0) MMA(c1) MMA(c2) MMA(c3) }            // independent
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (c0[0] + c1[0] + c2[0] + c3[0] + a[0] == 12345.f) out[0] = 0;
}
int main() {
  long long* d; cudaMalloc(&d, 8 * 256);
  const int iters = 4096;
  for (int warps : {4, 8, 16}) for (int mode : {1, 2, 6, 3, 7}) {
    k<<<148, warps * 32>>>(iters, d, mode); cudaDeviceSynchronize();
    k<<<148, warps * 32>>>(iters, d, mode); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d mode %d (%s%s%s): %.1f cycles/iter (4 mma + 1 ldmatrix per iter where enabled)\n", warps, mode,
           mode & 1 ? "ldmatrix " : "", mode & 2 ? "mma " : "", mode & 4 ? "dependent" : "", (double)h / iters);
  }
  return 0;
}
