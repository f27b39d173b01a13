// Measurement-only: hardware warp slots handed to co-resident CTAs of 1, 2 and 4 warps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(unsigned *rec, int hold) {
    unsigned hw, sm; asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw)); asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    if ((threadIdx.x & 31) == 0) { unsigned i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rec[2 * i] = sm; rec[2 * i + 1] = hw; }
    long long t0 = clock64(); while (clock64() - t0 < hold) { }   // keep the CTAs co-resident
}
int main() {
    unsigned *d, h[2 * 148 * 8 * 4];
    cudaMalloc(&d, sizeof h);
    for (int warps : {1, 2, 4}) {
        int grid = 148 * 7;
        cudaMemset(d, 0xff, sizeof h);
        probe<<<grid, 32 * warps>>>(d, 2000000); cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(unsigned) * 2 * grid * warps, cudaMemcpyDeviceToHost);
        printf("CTAs of %d warp(s), 7 per SM; hardware warp slots of the CTAs resident on SM %u:", warps, h[0]);
        for (int i = 0; i < grid * warps; ++i) if (h[2 * i] == h[0]) printf(" %s%u", (i % warps == 0) ? "|" : "", h[2 * i + 1]);
        printf("\n");
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
