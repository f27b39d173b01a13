// Measurement-only: cost of a CTA barrier and of concurrent REDUX / LDS with 8 warps (the SGD team kernel's
// per-row synchronisation).  nvcc -arch=sm_100a -o barrier_probe barrier_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int KIND>
__global__ void probe(float *out, long long *cyc, float seed) {
    __shared__ float sh[512];
    float v = seed + threadIdx.x * 1e-3f;
    int iv = threadIdx.x + 7;
    sh[threadIdx.x] = v; sh[threadIdx.x + 256 & 511] = v;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (KIND == 0) { __syncthreads(); }
        if (KIND == 1) { v = fmaf(v, 1.0001f, 0.5f); __syncthreads(); }
        if (KIND == 2) { iv = __reduce_add_sync(0xffffffffu, iv) + 1; }                       // REDUX, all warps at once
        if (KIND == 3) { sh[threadIdx.x] = v; __syncthreads(); v += sh[(threadIdx.x + 32) & 255]; __syncthreads(); }  // STS, BAR, LDS, BAR
        if (KIND == 4) { sh[threadIdx.x] = v; __syncthreads(); v += sh[(threadIdx.x + 32) & 255]; }                   // STS, BAR, LDS (double buffer not needed for timing)
        if (KIND == 5) { asm volatile("bar.sync 1, 256;"); }
        if (KIND == 6) { asm volatile("bar.arrive 1, 256;"); asm volatile("bar.sync 2, 256;"); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v + iv;
}
int main() {
    float *d; long long *c, h;
    cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    const char *names[] = {"BAR.SYNC (256 thr)", "FFMA + BAR.SYNC", "REDUX x 8 warps concurrently", "STS+BAR+LDS+BAR", "STS+BAR+LDS", "bar.sync 1,256", "bar.arrive + bar.sync(other)"};
#define RUN(K, T) probe<K><<<1, T>>>(d, c, 0.7f); probe<K><<<1, T>>>(d, c, 0.7f); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-34s %4d threads %6.1f cycles / iteration\n", names[K], T, (double)h / N);
    RUN(0, 256) RUN(0, 128) RUN(0, 64) RUN(1, 256) RUN(2, 256) RUN(2, 32) RUN(3, 256) RUN(4, 256) RUN(5, 256)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
