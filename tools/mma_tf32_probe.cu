// Issue rate of the legacy warp-level mma.sync.m16n8k8 tf32 on sm_100a (for moving the likelihood pass's
// pre-activations X.W1 off the FMA pipe: 3xTF32 needs three of these per 16 rows x 8 hidden units).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_tf32_probe tools/mma_tf32_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ILP>
__global__ void mma_loop(float *out, int iters) {
    uint32_t a[4], b[2];
    for (int k = 0; k < 4; ++k) a[k] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + k);
    for (int k = 0; k < 2; ++k) b[k] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + k);
    float c[ILP][4];
    for (int q = 0; q < ILP; ++q) for (int k = 0; k < 4; ++k) c[q][k] = q + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int q = 0; q < ILP; ++q)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[q][0]), "+f"(c[q][1]), "+f"(c[q][2]), "+f"(c[q][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    float s = 0.f;
    for (int q = 0; q < ILP; ++q) for (int k = 0; k < 4; ++k) s += c[q][k];
    if (s == 12345.678f) out[0] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 2048, threads = 256;
    float *d;
    cudaMalloc(&d, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bps = 1; bps <= 8; bps *= 2) {
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            mma_loop<8><<<sms * bps, threads>>>(d, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r && ms < best) best = ms;
        }
        const double inst = (double)sms * bps * threads / 32 * iters * 8;
        printf("%d warps/SMSP: %.3f ms, %.2f cycles per mma.sync per SMSP, %.1f TFLOP/s tf32\n", bps * threads / 32 / 4, best,
               best * 1e-3 * p.clockRate * 1e3 / (inst / (sms * 4)), inst * 16 * 8 * 8 * 2 / (best * 1e-3) / 1e12);
    }
    return 0;
}
