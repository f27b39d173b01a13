// Measurement-only: throughput of independent REDUX.SUM per warp (the row-view output layer of the SGD team
// would issue 10 per warp and row).  nvcc -arch=sm_100a -o redux_probe redux_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int NR>
__global__ void probe(int *out, long long *cyc) {
    int v[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) v[k] = threadIdx.x * (k + 3) + 7;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        int s[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) s[k] = __reduce_add_sync(0xffffffffu, v[k]);
#pragma unroll
        for (int k = 0; k < NR; ++k) v[k] = v[k] * 3 + (s[k] & 15);     // dependent on the result: next round waits for all
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    int a = 0;
#pragma unroll
    for (int k = 0; k < NR; ++k) a += v[k];
    out[threadIdx.x] = a;
}
int main() {
    int *d; long long *c, h;
    cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
#define RUN(NR, T) probe<NR><<<1, T>>>(d, c); probe<NR><<<1, T>>>(d, c); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%2d independent REDUX per round, %3d threads: %6.1f cycles / round\n", NR, T, (double)h / N);
    RUN(1, 32) RUN(2, 32) RUN(5, 32) RUN(10, 32) RUN(1, 256) RUN(2, 256) RUN(5, 256) RUN(10, 256)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
