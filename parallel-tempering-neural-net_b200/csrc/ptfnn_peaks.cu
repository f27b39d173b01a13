// Measurement-only micro-benchmarks (NOT part of the product ABI): the on-chip peaks the chain
// kernel is bound by -- FP32 FMA issue, MUFU (ex2/rcp) issue, shared-memory broadcast bandwidth.
// MEASURED_PEAKS.json only has HBM and bf16 tensor numbers; SURVEY 8(d) asks the builder to
// measure these on the box.  Built as csrc/libptfnn_peaks.so, used by bench.py only.
#include <cuda_runtime.h>
#include <stdint.h>

template <int ILP>
__global__ void fma_kernel(float *out, int iters, float a, float b) {
    float v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = threadIdx.x * 1e-3f + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) v[k] = fmaf(v[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k];
    if (s == 12345.678f) out[0] = s;
}

// ex2 -> +1 -> rcp chains (the sigmoid core).  The FADD between the two MUFU ops keeps ptxas from
// folding rcp(ex2(x)) into ex2(-x), which an earlier version of this probe suffered from (it read
// 2x the real rate); tools/contention_probe.cu confirms 16 MUFU lanes/clk/SM on B200.
template <int ILP>
__global__ void mufu_kernel(float *out, int iters) {
    float v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = 0.5f + threadIdx.x * 1e-4f + k * 1e-2f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            float e;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[k]));
            e += 1.0f;
            asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[k]) : "f"(e));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k];
    if (s == 12345.678f) out[0] = s;
}

// three-register FFMA (no immediate / constant operand), the form real kernels issue
template <int ILP>
__global__ void fma3_kernel(float *out, int iters, const float *ab) {
    float v[ILP], a[ILP], b[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { v[k] = threadIdx.x * 1e-3f + k; a[k] = ab[k] + threadIdx.x * 1e-9f; b[k] = ab[ILP + k]; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) v[k] = fmaf(v[k], a[k], b[k]);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k];
    if (s == 12345.678f) out[0] = s;
}

__global__ void smem_kernel(float *out, int iters) {
    __shared__ float4 buf[1024];
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) buf[k] = make_float4(k, k + 1, k + 2, k + 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 v = buf[(idx + 32 * k) & 1023];   // conflict-free 16 B per lane
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx = (idx + 7) & 1023;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

static float time_ms(void (*launch)(int, float *), int sms, float *d) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch(sms, d);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        launch(sms, d);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

static const int kIters = 4096, kThreads = 256, kBlocksPerSm = 8;
static void launch_fma(int sms, float *d) { fma_kernel<8><<<sms * kBlocksPerSm, kThreads>>>(d, kIters, 1.0001f, 0.5f); }
static void launch_mufu(int sms, float *d) { mufu_kernel<8><<<sms * kBlocksPerSm, kThreads>>>(d, kIters / 4); }
static void launch_smem(int sms, float *d) { smem_kernel<<<sms * kBlocksPerSm, kThreads>>>(d, kIters / 4); }
static float *g_ab = nullptr;
static void launch_fma3(int sms, float *d) { fma3_kernel<8><<<sms * kBlocksPerSm, kThreads>>>(d, kIters, g_ab); }

// out[0] = FP32 FMA TFLOP/s (constant-operand form), out[1] = MUFU Gop/s (ex2+rcp counted as 2 ops),
// out[2] = shared-memory GB/s, out[3] = FP32 FMA TFLOP/s (three-register form)
extern "C" int ptfnn_measure_peaks(int device, double *out) {
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return -2;
    const int sms = p.multiProcessorCount;
    float *d;
    if (cudaMalloc(&d, 16) != cudaSuccess) return -2;
    const double threads = (double)sms * kBlocksPerSm * kThreads;
    float ms = time_ms(launch_fma, sms, d);
    out[0] = threads * kIters * 8 * 2.0 / (ms * 1e-3) / 1e12;
    ms = time_ms(launch_mufu, sms, d);
    out[1] = threads * (kIters / 4) * 8 * 2.0 / (ms * 1e-3) / 1e9;
    ms = time_ms(launch_smem, sms, d);
    out[2] = threads * (kIters / 4) * 8 * 16.0 / (ms * 1e-3) / 1e9;
    float hab[16];
    for (int k = 0; k < 16; ++k) hab[k] = k < 8 ? 1.0001f : 0.5f;
    cudaMalloc(&g_ab, sizeof hab);
    cudaMemcpy(g_ab, hab, sizeof hab, cudaMemcpyHostToDevice);
    ms = time_ms(launch_fma3, sms, d);
    out[3] = threads * kIters * 8 * 2.0 / (ms * 1e-3) / 1e12;
    cudaFree(g_ab);
    cudaFree(d);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
