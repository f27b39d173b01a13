// Throughput of three-register FFMA against packed FFMA2 (f32x2) on sm_100a: does packing two FMAs into one
// instruction raise the FMA rate of a throughput-bound kernel, or only save issue slots?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void fma3(float *out, int iters, const float *ab) {
    float v[ILP], a[ILP], b[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { v[k] = threadIdx.x * 1e-3f + k; a[k] = ab[k] + threadIdx.x * 1e-9f; b[k] = ab[ILP + k]; }
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < ILP; ++k) v[k] = fmaf(v[k], a[k], b[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k];
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void fma2(float *out, int iters, const float *ab) {
    float2 v[ILP], a[ILP], b[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
        v[k] = make_float2(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f + k);
        a[k] = make_float2(ab[k] + threadIdx.x * 1e-9f, ab[k] + threadIdx.x * 2e-9f);
        b[k] = make_float2(ab[ILP + k], ab[ILP + k] * 0.5f);
    }
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < ILP; ++k) v[k] = __ffma2_rn(v[k], a[k], b[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k].x + v[k].y;
    if (s == 12345.678f) out[0] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 4096, threads = 256, bps = 8;
    float *d, *ab, hab[16];
    for (int k = 0; k < 16; ++k) hab[k] = k < 8 ? 1.0001f : 0.5f;
    cudaMalloc(&d, 16); cudaMalloc(&ab, sizeof hab);
    cudaMemcpy(ab, hab, sizeof hab, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int r = 0; r < 6; ++r) {
            cudaEventRecord(e0);
            if (which == 0) fma3<8><<<sms * bps, threads>>>(d, iters, ab);
            else fma2<8><<<sms * bps, threads>>>(d, iters, ab);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        const double inst = (double)sms * bps * threads / 32 * iters * 8;           // warp instructions
        const double clk = p.clockRate * 1e3;
        printf("%s: %.3f ms, %.1f TFLOP/s, %.2f cycles per warp instruction per SMSP\n", which ? "FFMA2 (3 register pairs)" : "FFMA  (3 registers)     ",
               best, inst * 32 * (which ? 4.0 : 2.0) / (best * 1e-3) / 1e12, best * 1e-3 * clk / (inst / (sms * 4)));
    }
    return 0;
}
