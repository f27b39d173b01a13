// Measurement-only: per-phase clock stamps of the wide-hidden SGD team kernel (16-256-10).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPTFNN_TEAM_TRACE -o team_trace team_trace.cu
#include <cstdio>
#include <vector>
#include <cstdlib>
#include "../parallel-tempering-neural-net_b200/csrc/ptfnn_kernels.cuh"
using namespace ptfnn;
int main() {
    constexpr int I = 16, H = 256, O = 10, NT = 256, P = I * H + H * O + H + O, IP = 16, N = 2048;
    std::vector<float> x((size_t)N * IP), y(N), w(P);
    srand(1);
    for (auto &v : x) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto &v : y) v = (float)(rand() % 10);
    for (auto &v : w) v = (rand() / (float)RAND_MAX - 0.5f) * 0.6f;
    float *dx, *dy, *dw, *dout;
    cudaMalloc(&dx, x.size() * 4); cudaMalloc(&dy, y.size() * 4); cudaMalloc(&dw, P * 4); cudaMalloc(&dout, P * 4);
    cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), y.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), P * 4, cudaMemcpyHostToDevice);
    const size_t smem = (((size_t)P * 4 + 15) & ~(size_t)15) + 16 + (size_t)2 * (kTileRows * IP * 4 + kTileRows * 4) + (size_t)team_smem_floats(H, O) * 4 + 16;
    auto k = op_sgd_kernel<I, H, O, kTaskCls, NT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    DataView d{dx, dy, N};
    for (int rep = 0; rep < 2; ++rep) k<<<1, NT, smem>>>(dw, dout, d, 0.01f, 1);
    cudaDeviceSynchronize();
    static long long t[8][16][8];
    cudaMemcpyFromSymbol(t, g_team_trace, sizeof t);
    printf("err=%s\nstamps: 0 row start | 1 hid published | 2 deferred updates done | 3 all hid visible | 4 od published | 5 column update + look-ahead done | 6 all od visible | 7 row end\n", cudaGetErrorString(cudaGetLastError()));
    for (int wv : {0, 3, 7})
        for (int r = 4; r < 8; ++r) {
            printf("warp %d row %d:", wv, 64 + r);
            for (int s = 1; s < 8; ++s) printf(" %5lld", t[wv][r][s] - t[wv][r][s - 1]);
            printf("  | row total %5lld\n", t[wv][r + 1][0] - t[wv][r][0]);
        }
    return 0;
}
