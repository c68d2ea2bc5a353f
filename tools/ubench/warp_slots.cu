// Which hardware warp slot (%warpid; scheduler = %warpid % 4) does each warp of a CTA get when several
// CTAs share an SM?  Prints, per SM-resident CTA, the slot of every warp.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/warp_slots tools/ubench/warp_slots.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out, int regs_dummy) {
  unsigned wid, smid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  // keep the CTA resident long enough for the co-resident CTA to arrive
  long long t0 = clock64();
  while (clock64() - t0 < 200000) {}
  if ((threadIdx.x & 31) == 0) {
    int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    out[(blockIdx.x * nw + w) * 2] = smid;
    out[(blockIdx.x * nw + w) * 2 + 1] = wid;
  }
}
int main() {
  int* d; cudaMalloc(&d, 1 << 20);
  static int h[1 << 18];
  for (int nw : {10, 12, 19, 20}) {
    int grid = nw >= 19 ? 148 : 296;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<grid, nw * 32, nw >= 19 ? 0 : 100 * 1024>>>(d, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, grid * nw * 8, cudaMemcpyDeviceToHost);
    printf("warps/CTA %d:\n", nw);
    // print CTAs living on the SM of CTA 0 and CTA 1
    for (int target = 0; target < 2; ++target) {
      int sm = h[(target * nw) * 2];
      for (int b = 0; b < grid; ++b) if (h[(b * nw) * 2] == sm) {
        printf("  SM %3d CTA %3d slots:", sm, b);
        for (int w = 0; w < nw; ++w) printf(" %2d", h[(b * nw + w) * 2 + 1]);
        printf("   sched:");
        for (int w = 0; w < nw; ++w) printf(" %d", h[(b * nw + w) * 2 + 1] % 4);
        printf("\n");
      }
    }
  }
  return 0;
}
