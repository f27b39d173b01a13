// Measurement-only: dependent-issue latency (cycles) of the instructions on the critical path of the
// serial SGD recurrence, on the box's GPU.  nvcc -arch=sm_100a -o latency_probe latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define N 4096
template <int KIND>
__global__ void probe(float *out, long long *cyc, float seed) {
    float v = seed + threadIdx.x * 1e-3f;
    int iv = threadIdx.x + 7;
    float v2 = v * 0.5f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (KIND == 0) v = fmaf(v, 1.0001f, 0.5f);
        if (KIND == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v));
        if (KIND == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v));
        if (KIND == 3) v += __shfl_xor_sync(0xffffffffu, v, 1 << (i % 5));
        if (KIND == 4) { iv = __reduce_add_sync(0xffffffffu, iv) + 1; }
        if (KIND == 5) { iv = __float2int_rn(v); v = (float)iv * 0.999f; }          // F2I + I2F + FMUL
        if (KIND == 6) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v)); v += 1.0f; asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v)); } // sigmoid core
        if (KIND == 7) { float t = __shfl_xor_sync(0xffffffffu, v, 16); v = v + t * 1e-9f; }       // shfl + ffma
        if (KIND == 8) { iv = __float2int_rn(v * 1048576.0f); iv = __reduce_add_sync(0xffffffffu, iv); v = (float)iv * (1.0f / 33554432.0f); } // fixed-point allreduce
        if (KIND == 9) { unsigned b = __ballot_sync(0xffffffffu, v > 0.5f); v += (float)(b & 1) * 1e-9f; }
        if (KIND == 10) { float2 a = make_float2(v, v2), b = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f); a = __ffma2_rn(a, b, c); v = a.x; v2 = a.y; }   // FFMA2 chain
        if (KIND == 11) { iv = __float2int_rn(v); v = __int_as_float((iv & 0xff) | 0x3f800000); }           // F2I + LOP3
        if (KIND == 12) { v = (float)iv; iv = __float_as_int(v) & 0xffff; }                                    // I2F + LOP3
        if (KIND == 13) { iv = __reduce_add_sync(0xffffffffu, iv); v = (float)iv * 1e-3f; iv = __float_as_int(v) & 0xfff; }  // REDUX + I2F + FMUL + LOP3
        if (KIND == 14) { float t = v + 12582912.0f; iv = __float_as_int(t); iv = __reduce_add_sync(0xffffffffu, iv); v = __int_as_float((iv & 0x3fffff) | 0x3f000000); } // magic FADD + REDUX + LOP3
        if (KIND == 15) { v = fmaf(v, 1.0001f, 0.5f); v = fmaxf(v, 0.1f); }                                    // FFMA + FMNMX (cross pipe)
        if (KIND == 16) { float t = __shfl_sync(0xffffffffu, v, i & 31); v = v + t * 1e-9f; }                 // SHFL.IDX + FFMA
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
    out[threadIdx.x] = v + iv + v2;
}

int main() {
    float *d; long long *c, h;
    cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
    const char *names[] = {"FFMA", "MUFU.EX2", "MUFU.RCP", "SHFL.BFLY+FADD", "REDUX.SUM.S32+IADD", "F2I+I2F+FMUL",
                           "EX2+FADD+RCP", "SHFL+FFMA", "FMUL+F2I+REDUX+I2F+FMUL", "VOTE.BALLOT+...",
                           "FFMA2", "F2I+LOP3", "I2F+LOP3", "REDUX+I2F+FMUL+LOP3", "FADD(magic)+REDUX+LOP3", "FFMA+FMNMX", "SHFL.IDX+FFMA"};
#define RUN(K) probe<K><<<1, 32>>>(d, c, 0.7f); probe<K><<<1, 32>>>(d, c, 0.7f); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-28s %6.1f cycles / iteration\n", names[K], (double)h / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
