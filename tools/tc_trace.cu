// Measurement only: per-phase clock stamps of the warp-specialised tcgen05 likelihood pass (K5), one CTA, first tiles.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPTFNN_TC_TRACE -I../parallel-tempering-neural-net_b200/csrc -o tc_trace tc_trace.cu
#include <cstdio>
#include <vector>
#include "ptfnn_kernels.cuh"
using namespace ptfnn;
int main() {
    constexpr int I = 16, H = 256, O = 10, NT = tc::kThreads, P = I * H + H * O + H + O, N = 20000, IP = 16;
    std::vector<float> x((size_t)N * IP), y(N + 4, 0.f), w(P);
    unsigned s = 7u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f - 0.5f; };
    for (auto &v : x) v = 2.f * rnd();
    for (auto &v : w) v = 0.6f * rnd();
    for (int r = 0; r < N; ++r) y[r] = (float)(r % O);
    float *dx, *dy, *dw, *dt; double *ds;
    const size_t ntiles = (N + 127) / 128, tile_f = tc::a_tile_floats(I);
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dy, y.size() * 4); cudaMalloc(&dw, w.size() * 4); cudaMalloc(&dt, ntiles * tile_f * 4); cudaMalloc(&ds, 64);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), y.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    tc::pack_a_kernel<I><<<256, 256>>>(dx, N, IP, dt);
    auto k = op_forward_tc_kernel<I, H, O, 1, NT>;
    const int smem = tc::Smem<I, H, O>::total;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 296}) {
        double *dsg; cudaMalloc(&dsg, 24 * grid);
        float *dwb; cudaMalloc(&dwb, (size_t)P * 4 * grid);
        float *dfx; cudaMalloc(&dfx, (size_t)N * 4 * grid);
        for (int g = 0; g < grid; ++g) cudaMemcpy(dwb + (size_t)g * P, dw, P * 4, cudaMemcpyDeviceToDevice);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<<<grid, NT, smem>>>(dwb, dt, dy, N, dfx, nullptr, dsg);
        cudaEventRecord(a);
        k<<<grid, NT, smem>>>(dwb, dt, dy, N, dfx, nullptr, dsg);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("grid %3d: %.3f ms = %.0f cycles per 128-row tile and CTA (%s)\n", grid, ms, ms * 1e-3 * 1.965e9 / ntiles, cudaGetErrorString(e));
        {
            static long long tr[2][8][64];
            cudaMemcpyFromSymbol(tr, tc::g_tc_trace, sizeof tr);
            for (int t = 2; t < 5; ++t) {
                const long long t0 = tr[0][t][40];
                printf("tile %d epilogue warp 0 (cycles from its Z wait): Z0 ready %lld", t, tr[0][t][41] - t0);
                for (int g = 0; g < 8; ++g) {
                    if (g == 4) printf("\n      Z1 wait %lld -> ready %lld", tr[0][t][42] - t0, tr[0][t][43] - t0);
                    printf("\n      sb %d: ld done %lld, sigmoids done %lld, L free %lld, arrived %lld", g, tr[0][t][4 * g] - t0, tr[0][t][4 * g + 1] - t0, tr[0][t][4 * g + 2] - t0, tr[0][t][4 * g + 3] - t0);
                }
                printf("\n      D wait %lld -> ready %lld; next tile's Z wait at %lld\n", tr[0][t][44] - t0, tr[0][t][45] - t0, tr[0][t + 1][40] - t0);
                printf("tile %d MMA warp (same origin): A wait %lld -> %lld, layer 1 issued %lld", t, tr[1][t][0] - t0, tr[1][t][1] - t0, tr[1][t][2] - t0);
                for (int g = 0; g < 8; ++g) printf("\n      sb %d: H ready %lld, MMAs issued + committed %lld", g, tr[1][t][3 + 2 * g] - t0, tr[1][t][4 + 2 * g] - t0);
                printf("\n");
            }
        }
    }
    return 0;
}
