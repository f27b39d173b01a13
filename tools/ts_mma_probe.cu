// Measurement / validation only: layer 2 of the wide-hidden likelihood pass on tcgen05 with the A operand
// (hidden activations) in TENSOR MEMORY.  Checks, in isolation, everything the K5 pass relies on:
//   * tcgen05.st 32x32b.x32 places A[row = lane][k = column] where a TS-mode MMA expects it,
//   * tcgen05.mma kind::tf32 with A from TMEM, M = 128, N = 16, K = 8 per instruction, B from shared memory,
//   * a region written by an SS-mode MMA (an accumulator) can be overwritten in place and used as A,
//   * 3xTF32 (hi/lo split of both operands) reaches fp32-class accuracy,
//   * cycles per N = 16 / N = 32 / N = 256 instruction (tensor-pipe pacing for small N).
// nvcc -gencode arch=compute_100a,code=sm_100a -I../parallel-tempering-neural-net_b200/csrc -o ts_mma_probe ts_mma_probe.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include <algorithm>
#include "ptfnn_kernels.cuh"

using namespace ptfnn;
using namespace ptfnn::tc;

constexpr int K2 = 64;      // hidden units of the probe (K of layer 2)
constexpr int N2 = 16;      // padded outputs

__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// hid: [128][K2] activations in (0,1); w2: [K2][N2]; out: [128][N2]; mode 0 = 3xTF32, 1 = plain tf32
__global__ void __launch_bounds__(128) probe(const float *hid, const float *w2, float *out, int mode, long long *cyc) {
    __shared__ __align__(128) unsigned char s_b[2 * (K2 / 4) * N2 * 16];     // hi | lo, chunk c (4 k-values) x row n x 16 B
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int BCH = N2 * 16, BHALF = (K2 / 4) * BCH;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (tid < 32) tmem_alloc(&slot, 256);
    for (int idx = tid; idx < (K2 / 4) * N2; idx += 128) {
        const int c = idx / N2, n = idx % N2;
        float hi[4], lo[4];
        for (int e = 0; e < 4; ++e) {
            const float v = w2[(c * 4 + e) * N2 + n];
            hi[e] = tf32_hi(v); lo[e] = tf32_hi(v - hi[e]);
        }
        *reinterpret_cast<float4 *>(s_b + c * BCH + n * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(s_b + BHALF + c * BCH + n * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t col_hi = 0, col_lo = K2, col_d = 2 * K2;          // TMEM columns: hid_hi | hid_lo | D (16)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // ---- A operand: row = tid; written straight from registers
    for (int cb = 0; cb < K2 / 32; ++cb) {
        float hi[32], lo[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float v = hid[tid * K2 + cb * 32 + k];
            hi[k] = tf32_hi(v); lo[k] = tf32_hi(v - hi[k]);
        }
        tmem_st32(tm + lane_base + col_hi + cb * 32, hi);
        tmem_st32(tm + lane_base + col_lo + cb * 32, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    const uint32_t b_base = smem_u32(s_b);
    constexpr uint32_t IDESC = instr_desc_tf32(128, N2);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        tc_fence_after();
        t0 = clock64();
        bool acc = false;
        for (int ks = 0; ks < K2 / 8; ++ks) {
            const uint64_t bhi = smem_desc(b_base + ks * 2 * BCH, BCH, 128), blo = smem_desc(b_base + BHALF + ks * 2 * BCH, BCH, 128);
            mma_tf32_ts(tm + col_d, tm + col_hi + ks * 8, bhi, IDESC, acc);
            acc = true;
            if (mode == 0) {
                mma_tf32_ts(tm + col_d, tm + col_lo + ks * 8, bhi, IDESC, true);
                mma_tf32_ts(tm + col_d, tm + col_hi + ks * 8, blo, IDESC, true);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
    tc_fence_after();
    float d[16];
    tmem_ld16(tm + lane_base + col_d, d);
    for (int n = 0; n < N2; ++n) out[tid * N2 + n] = d[n];
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, 256);
}

// pacing: `reps` back-to-back TS MMAs of width N, round-robin over ND independent accumulators; the grid may put
// two CTAs on every SM (do their MMAs overlap, or is the cost per instruction pipe occupancy?)
template <int N, int ND>
__global__ void __launch_bounds__(128, 2) pace(long long *cyc, int reps) {
    __shared__ __align__(128) unsigned char s_b[2 * 256 * 16];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (tid < 32) tmem_alloc(&slot, 256);
    for (int i = tid; i < 2 * 256 * 4; i += 128) reinterpret_cast<float *>(s_b)[i] = 0.0f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (tid == 0) {
        const uint64_t bd = smem_desc(smem_u32(s_b), N * 16, 128);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) mma_tf32_ts(tm + 64 + (r % ND) * (N < 32 ? 32 : N), tm + (r & 7) * 8, bd, instr_desc_tf32(128, N), r >= ND);
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        cyc[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, 256);
}

int main() {
    std::vector<float> hid(128 * K2), w2(K2 * N2), out(128 * N2);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
    for (auto &v : hid) v = rnd();
    for (auto &v : w2) v = 4.0f * rnd() - 2.0f;
    float *dh, *dw, *dout; long long *dc, hc;
    cudaMalloc(&dh, hid.size() * 4); cudaMalloc(&dw, w2.size() * 4); cudaMalloc(&dout, out.size() * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dh, hid.data(), hid.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode) {
        probe<<<1, 128>>>(dh, dw, dout, mode, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("probe mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
        double worst = 0.0;
        int wr = 0, wn = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N2; ++n) {
                double ref = 0.0;
                for (int k = 0; k < K2; ++k) ref += (double)hid[r * K2 + k] * (double)w2[k * N2 + n];
                const double err = std::fabs(ref - out[r * N2 + n]);
                if (err > worst) { worst = err; wr = r; wn = n; }
            }
        printf("TS-mode layer 2 (%s): max abs err %.3e at row %d col %d (values ~ %.1f); %lld cycles for %d MMAs incl. commit + wait\n",
               mode == 0 ? "3xTF32" : "plain tf32", worst, wr, wn, std::sqrt((double)K2) * 0.6, hc, (K2 / 8) * (mode == 0 ? 3 : 1));
    }
    const int reps = 256;
    long long *dcs; cudaMalloc(&dcs, 8 * 512);
    std::vector<long long> hcs(512);
#define PACE(N, ND, GRID) pace<N, ND><<<GRID, 128>>>(dcs, reps); cudaDeviceSynchronize(); cudaMemcpy(hcs.data(), dcs, 8 * GRID, cudaMemcpyDeviceToHost); \
    { double mx = 0; for (int i = 0; i < GRID; ++i) mx = std::max(mx, (double)hcs[i]); \
      printf("TS MMA M=128 N=%3d K=8, %d accumulators, %3d CTAs: %.1f cycles per instruction and CTA (%d back to back)\n", N, ND, GRID, mx / reps, reps); }
    PACE(16, 1, 1) PACE(16, 2, 1) PACE(16, 4, 1) PACE(32, 1, 1) PACE(32, 2, 1) PACE(64, 1, 1) PACE(64, 2, 1) PACE(128, 1, 1) PACE(192, 1, 1)
    PACE(16, 1, 296) PACE(16, 2, 296) PACE(32, 1, 296) PACE(64, 1, 296) PACE(128, 1, 296)
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
