// Measurement-only: does the occupancy calculator (and the hardware) co-schedule two CTAs that allocate TMEM?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int USE_TMEM, int COLS>
__global__ void __launch_bounds__(256, 2) k(int *out, long long spin) {
    __shared__ uint32_t slot;
    uint32_t base = 0;
    if (USE_TMEM) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __syncthreads();
        base = slot;
    }
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) out[blockIdx.x] = (int)smid;
    __syncthreads();
    if (USE_TMEM && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS) : "memory");
}
template <class K> void run(const char *name, K kern) {
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    int *d; cudaMalloc(&d, 4096 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); kern<<<296, 256>>>(d, 2000000); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s regs %3d occupancy API %d; 296 CTAs x 1.0 ms spin took %.2f ms (%s)  err=%s\n", name, fa.numRegs, occ, ms, ms < 1.6 ? "2 per SM co-resident" : "1 per SM", cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run("no TMEM", k<0, 256>);
    run("TMEM 256 columns", k<1, 256>);
    run("TMEM 128 columns", k<1, 128>);
    return 0;
}
