// Measurement-only: how the dependent-chain cost of SHFL / REDUX / MUFU / F2I grows when K warps of
// one SM run the same chain concurrently (is the unit shared per SM, per sub-partition, pipelined?).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int KIND>
__global__ void probe(float *out, long long *cyc, float seed) {
    float v = seed + threadIdx.x * 1e-3f;
    int iv = threadIdx.x + 7;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (KIND == 0) v += __shfl_xor_sync(0xffffffffu, v, 1 << (i % 5));
        if (KIND == 1) { iv = __reduce_add_sync(0xffffffffu, iv) + 1; }
        if (KIND == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v)); v += 1.0f; asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v)); }
        if (KIND == 3) { iv = __float2int_rn(v); v = (float)iv * 0.999f; }
        if (KIND == 4) { unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(v) & 0x7fffffffu); iv = __float2int_rn(v * 4194304.0f); iv = __reduce_add_sync(0xffffffffu, iv); v = (m < 0x417e6666u) ? (float)iv * (1.0f / 134217728.0f) : v; }
        if (KIND == 5) { iv = __float2int_rn(v * 4194304.0f); iv = __reduce_add_sync(0xffffffffu, iv); v = (float)iv * (1.0f / 134217728.0f); }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long *)cyc, (unsigned long long)(t1 - t0));
    out[threadIdx.x] = v + iv;
}
int main() {
    float *d; long long *c, h;
    cudaMalloc(&d, 1 << 16); cudaMalloc(&c, 8);
    const char *names[] = {"SHFL+FADD", "REDUX.ADD+IADD", "EX2+FADD+RCP", "F2I+I2F+FMUL", "guarded fixed allreduce", "fixed allreduce"};
    int ks[] = {1, 2, 4, 8, 16, 32};
    printf("%-26s", "warps per SM (1 CTA):");
    for (int k : ks) printf("%8d", k);
    printf("\n");
#define RUN(K) printf("%-26s", names[K]); for (int k : ks) { cudaMemset(c, 0, 8); probe<K><<<1, 32 * k>>>(d, c, 0.7f); cudaDeviceSynchronize(); cudaMemset(c, 0, 8); probe<K><<<1, 32 * k>>>(d, c, 0.7f); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%8.1f", (double)h / N); } printf("\n");
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
