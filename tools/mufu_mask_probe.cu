// Measurement-only: does a MUFU / SHFL warp-instruction with few active lanes occupy the XU pipe for
// less time?  One warp, ILP independent ex2->add->rcp chains, `active` lanes enabled.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int ILP>
__global__ void probe(float *out, long long *cyc, int active) {
    float v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = 0.3f + threadIdx.x * 1e-3f + k * 0.01f;
    long long t0 = 0, t1 = 0;
    if ((int)threadIdx.x < active) {
        t0 = clock64();
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                float e;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[k]));
                e += 1.0f;
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(v[k]) : "f"(e));
            }
        }
        t1 = clock64();
    }
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    float s = 0; for (int k = 0; k < ILP; ++k) s += v[k];
    out[threadIdx.x] = s;
}
// independent broadcasts (all-gather by H shuffles) followed by a local dot product
template <int H>
__global__ void gather_probe(float *out, long long *cyc) {
    float v = 0.3f + threadIdx.x * 1e-3f;
    float w[H];
    for (int k = 0; k < H; ++k) w[k] = 0.01f * (k + 1);
    long long t0 = clock64();
#pragma unroll 2
    for (int i = 0; i < N; ++i) {
        float g[H];
#pragma unroll
        for (int k = 0; k < H; ++k) g[k] = __shfl_sync(0xffffffffu, v, k);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k + 1 < H; k += 2) { a = fmaf(g[k], w[k], a); b = fmaf(g[k + 1], w[k + 1], b); }
        if (H & 1) a = fmaf(g[H - 1], w[H - 1], a);
        v = (a + b) * 0.999f + 0.001f;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v;
}
int main() {
    float *d; long long *c, h;
    cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    int act[] = {32, 16, 8, 4, 1};
    printf("sigmoid core (EX2+FADD+RCP), cycles per iteration of ILP chains, one warp\n%-8s", "ILP");
    for (int a : act) printf("  act=%-3d", a);
    printf("\n");
#define RUN(I) printf("%-8d", I); for (int a : act) { probe<I><<<1, 32>>>(d, c, a); probe<I><<<1, 32>>>(d, c, a); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("  %7.1f", (double)h / N); } printf("\n");
    RUN(1) RUN(2) RUN(5) RUN(6) RUN(12)
#define RUNG(H) gather_probe<H><<<1, 32>>>(d, c); gather_probe<H><<<1, 32>>>(d, c); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("all-gather by %2d independent SHFL + dot + FFMA: %6.1f cycles\n", H, (double)h / N);
    RUNG(5) RUNG(10) RUNG(12) RUNG(16)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
