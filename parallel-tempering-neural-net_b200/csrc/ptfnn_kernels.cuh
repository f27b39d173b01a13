// Kernels of the sm_100a parallel-tempering FNN sampler (templated on the topology [I,H,O] and task).
//
//   K1  lik_rows      row-parallel forward + Gaussian / softmax-of-sigmoid reduction
//                     (Network.evaluate_proposal R:120-134 / C:134-153, likelihood_func R:200-205 / C:209-222)
//   K2  SgdWarp       the serial online-SGD recurrence ("Langevin gradient", R:99-118 + R:51-78 / C:72-82)
//   K3  chain_kernel  persistent per-temperature MCMC loop (ptReplica.run R:313-437 / C:313-448)
//   K4  swap round    grid barrier + sequential sweep + row exchange (R:427-437, R:659-690, R:741-752)
//
// One CTA per temperature.  Warp 0 owns the serial recurrence (hidden units live in its lanes'
// registers); the remaining warps run the row-parallel likelihood of the proposal at the same time.
#pragma once
#include "ptfnn_device.cuh"

namespace ptfnn {

template <int I>
struct IPad {
    static constexpr int value = (I + 3) & ~3;
};

struct DataView {
    const float *x;  // [n][IP] row-major, zero padded to a multiple of 4 floats (16-byte rows for TMA / LDS.128)
    const float *y;  // [n]     regression target (R:201) or class label as float (C:210)
    int n;
};

constexpr int kTileRows = 128;
constexpr int kSweepChunk = 128; // lhood fields staged per chunk by the in-kernel swap sweep
// Likelihood-pass sigmoids: MUFU ex2 + rcp (relative error ~4e-7, well inside the 1e-4 parity bar);
// log / exp of the softmax epilogue stay full precision.
constexpr bool kPreciseLik = false;  // rows per TMA tile when the training set is streamed instead of staged

// ==========================================================================================
// K2: serial SGD recurrence, one warp.  Lane l owns hidden units l, l+32, ... in registers.
//
// The recurrence is latency-bound (one dependent chain per row) and the warp issues in order, so
// both the LENGTH of the chain and the instruction COUNT per row matter.  Per row the chain is
//     EX2 -> FADD -> RCP            hidden sigmoid, pre-activation kept pre-multiplied by -log2(e)
//     -> FFMA(s) -> REDUX -> IADD   output-layer sum in block fixed point on the integer REDUX unit: the
//        lane's partial sum is accumulated ON TOP of the magic constant 1.5 * 2^23, so its float bit
//        pattern already is (constant + integer) and no F2I (22 cycles) is needed
//        (H <= 8: SHFL all-gather + local dot product with a replicated, pre-scaled W2 instead)
//     -> I2F -> FFMA                un-scale, subtract B2 and multiply by -log2(e) in ONE FFMA
//     -> EX2 -> FADD -> RCP         output sigmoid
//     -> {FADD | FFMA} -> FMUL      out_delta = (d - out) * (out - out^2)
//     -> FFMA                       next row's scaled pre-activation = stale value + out_delta * K
// where everything else (stale pre-activation of the next row with the not-yet-updated weights,
// K = -log2(e) lr (x_next.x + 1) hid (1 - hid) W2, the weight updates, the range guard of the fixed
// point sum) is independent of the chain and fills its MUFU / REDUX latencies.  With two hidden
// units per lane (H = 64) the lane-local arithmetic uses the packed FFMA2 / FMUL2 / FADD2
// instructions of sm_100 (two fp32 operations per issue slot).
// ==========================================================================================
constexpr float kL2E = 1.4426950408889634f;
constexpr float kMagic = 12582912.0f;            // 1.5 * 2^23: float(kMagic + v) has integer resolution for |v| < 2^22
constexpr int kMagicSum32 = 0x68000000;          // 32 * bits(kMagic) mod 2^32: what 32 lanes of bare magic add up to
constexpr int kGuardRows = 30;                   // rows covered by one choice of the fixed-point scale

// Packed pairs kept in 64-bit registers for the whole hidden-unit loop (building the pair from two
// scalar registers costs a MOV each time, which is what the loop is trying to get rid of).
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pack2(float lo, float hi) { f2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) { f2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// elementwise helpers over the HPL hidden units of a lane; adjacent pairs go through f32x2
template <int N>
__device__ __forceinline__ void vfma(float (&r)[N], const float (&a)[N], const float (&b)[N], const float (&c)[N]) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
        const float2 t = __ffma2_rn(make_float2(a[k], a[k + 1]), make_float2(b[k], b[k + 1]), make_float2(c[k], c[k + 1]));
        r[k] = t.x; r[k + 1] = t.y;
    }
    if (N & 1) r[N - 1] = fmaf(a[N - 1], b[N - 1], c[N - 1]);
}
template <int N>
__device__ __forceinline__ void vfma_s(float (&r)[N], const float (&a)[N], float s, const float (&c)[N]) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
        const float2 t = __ffma2_rn(make_float2(a[k], a[k + 1]), make_float2(s, s), make_float2(c[k], c[k + 1]));
        r[k] = t.x; r[k + 1] = t.y;
    }
    if (N & 1) r[N - 1] = fmaf(a[N - 1], s, c[N - 1]);
}
template <int N>
__device__ __forceinline__ void vmul(float (&r)[N], const float (&a)[N], const float (&b)[N]) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
        const float2 t = __fmul2_rn(make_float2(a[k], a[k + 1]), make_float2(b[k], b[k + 1]));
        r[k] = t.x; r[k + 1] = t.y;
    }
    if (N & 1) r[N - 1] = a[N - 1] * b[N - 1];
}
template <int N>
__device__ __forceinline__ void vmul_s(float (&r)[N], const float (&a)[N], float s) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
        const float2 t = __fmul2_rn(make_float2(a[k], a[k + 1]), make_float2(s, s));
        r[k] = t.x; r[k + 1] = t.y;
    }
    if (N & 1) r[N - 1] = a[N - 1] * s;
}
template <int N>
__device__ __forceinline__ void vadd_s(float (&r)[N], const float (&a)[N], float s) {
#pragma unroll
    for (int k = 0; k + 1 < N; k += 2) {
        const float2 t = __fadd2_rn(make_float2(a[k], a[k + 1]), make_float2(s, s));
        r[k] = t.x; r[k + 1] = t.y;
    }
    if (N & 1) r[N - 1] = a[N - 1] + s;
}

template <int I, int H, int O, int TASK>
struct SgdWarp {
    static constexpr int HPL = (H + 31) / 32;
    static constexpr int IP = IPad<I>::value;
    static constexpr int LEVELS = H > 16 ? 5 : H > 8 ? 4 : H > 4 ? 3 : H > 2 ? 2 : H > 1 ? 1 : 0;
    // Small H: the output layer is evaluated from an ALL-GATHER of the hidden activations -- H
    // independent SHFLs (pipelined: ~40 cycles for H = 5, measured) + a local dot product with a
    // replicated copy of W2 -- instead of a dependent reduction (REDUX path ~78 cycles).
    static constexpr bool GATHER = H <= 8;    // measured: H = 10 is faster on the REDUX path
    static constexpr int HG = GATHER ? H : 1;
    // hidden-unit index k is the FASTEST index so that pairs (k, k+1) feed the f32x2 instructions
    float w1[I][HPL], b1[HPL], w2[O][HPL], b2[O];
    float b2l[O];        // log2(e) * B2
    float w2q[O][HPL];   // REDUX path: fscale * W2 (a power of two: exact)
    float fscale, cdec;  // REDUX path: fixed-point scale 2^s and the decode factor -log2(e) 2^-s
    float w2f[HG][O];    // GATHER: full W2, identical in every lane (same update, bit for bit)

    // weight vector layout a1 (R:80-90): [W1 (I x H), W2 (H x O), B1 (H), B2 (O)]
    __device__ __forceinline__ void load(const float *w, int lane) {
#pragma unroll
        for (int k = 0; k < HPL; ++k) {
            const int h = lane + 32 * k;
            const bool a = h < H;
#pragma unroll
            for (int i = 0; i < I; ++i) w1[i][k] = a ? w[i * H + h] : 0.0f;
#pragma unroll
            for (int o = 0; o < O; ++o) w2[o][k] = a ? w[I * H + h * O + o] : 0.0f;
            // lanes without a hidden unit: z = x.0 - 1e4 -> sigmoid = 0 exactly (ex2 -> +inf, rcp -> 0),
            // so every one of their updates is an exact zero and no select sits on the dependent chain
            b1[k] = a ? w[I * H + H * O + h] : 1.0e4f;
        }
#pragma unroll
        for (int o = 0; o < O; ++o) { b2[o] = w[I * H + H * O + H + o]; b2l[o] = kL2E * b2[o]; }
        if constexpr (GATHER) {
#pragma unroll
            for (int h = 0; h < H; ++h)
#pragma unroll
                for (int o = 0; o < O; ++o) w2f[h][o] = w[I * H + h * O + o];
        }
        fscale = 1.0f; cdec = -kL2E;
        refresh(lane);
    }
    __device__ __forceinline__ void store(float *w, int lane) const {
#pragma unroll
        for (int k = 0; k < HPL; ++k) {
            const int h = lane + 32 * k;
            if (h < H) {
#pragma unroll
                for (int i = 0; i < I; ++i) w[i * H + h] = w1[i][k];
#pragma unroll
                for (int o = 0; o < O; ++o) w[I * H + h * O + o] = w2[o][k];
                w[I * H + H * O + h] = b1[k];
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int o = 0; o < O; ++o) w[I * H + H * O + H + o] = b2[o];
        }
    }
    // derived copy of the output layer for the next row; off the dependent chain
    __device__ __forceinline__ void refresh(int) {
        if constexpr (!GATHER) {
#pragma unroll
            for (int o = 0; o < O; ++o) vmul_s<HPL>(w2q[o], w2[o], fscale);
        }
    }
    // Chooses the fixed-point scale for the next `rows` rows (warp-uniform; one CREDUX, off the
    // chain).  |p_lane[o]| <= m = sum_k |W2[k][o]| because hid is in [0,1], and one row moves each
    // |W2| entry by at most lr * |out_delta| * hid <= lr / 4.  With 2^E > m the scale 2^(22-E) keeps
    // |p_lane * scale| < 2^22, where kMagic + v is exact to the integer: the sum is taken with a
    // resolution of 2^-22 relative to the binade of the largest lane bound (block fixed point, i.e.
    // fp32-class accuracy), and 32 lanes of < 2^22 cannot overflow the int32 REDUX.  Returns false
    // (use the SHFL butterfly) only for non-finite weights, so that NaN propagates as in the reference.
    __device__ __forceinline__ bool set_scale(int rows, float lr) {
        if constexpr (GATHER) return true;
        float m = 0.0f;
#pragma unroll
        for (int o = 0; o < O; ++o) {
            float a = 0.0f;
#pragma unroll
            for (int k = 0; k < HPL; ++k) a += fabsf(w2[o][k]);
            m = fmaxf(m, a);                                    // fmaxf drops NaN: checked separately below
            m = (a != a) ? __int_as_float(0x7fc00000) : m;
        }
        m = fmaxf(m + (float)(rows * HPL) * 0.25f * fabsf(lr), 9.765625e-4f);     // >= 2^-10: scale stays finite
        const unsigned int mb = __reduce_max_sync(0xffffffffu, __float_as_uint(m));   // NaN / Inf bits compare high
        const unsigned int ef = mb >> 23;                        // biased exponent of the bound, 2^(ef-127) <= m < 2^(ef-126)
        const bool ok = ef < 200u;                               // finite and far from overflow
        const unsigned int sf = ok ? 275u - ef : 127u;           // scale 2^(22 - (ef - 126))
        fscale = __uint_as_float(sf << 23);
        cdec = -kL2E * __uint_as_float((254u - sf) << 23);
        refresh(0);
        return ok;
    }

    // Scaled pre-activation of the hidden units for row x with the CURRENT weights:
    // zs = -log2(e) * (x.W1 - B1)   (bias is SUBTRACTED, R:52)
    __device__ __forceinline__ void preact(const float (&x)[IP], float (&zs)[HPL]) const {
        float t[HPL];
        vmul_s<HPL>(t, b1, -1.0f);
#pragma unroll
        for (int i = 0; i < I; ++i) vfma_s<HPL>(t, w1[i], x[i], t);
        vmul_s<HPL>(zs, t, -kL2E);
    }

    // One row: ForwardPass (R:51-55) then BackwardPass (R:57-78 / C:72-82).
    //   zs     in:  scaled pre-activations of THIS row (all earlier updates included)
    //          out: scaled pre-activations of the NEXT row xn = stale value (xn.W1_old - B1_old, computed
    //               while the output-layer reduction is in flight) + the effect of this row's rank-1
    //               update W1 += lr*hd (x) x, B1 -= lr*hd, i.e. lr*hd*(xn.x + 1), folded into one FFMA
    //               per output on the chain.
    //   FIX: the output-layer sum goes through the fixed-point REDUX (set_scale() returned true);
    //   a template parameter rather than a branch so that a row is ONE basic block the scheduler
    //   can interleave freely (no branch in the shadow of the REDUX).
    template <bool FIX>
    __device__ __forceinline__ void row(const float (&x)[IP], float yv, const float (&xn)[IP], float (&zs)[HPL],
                                        float lr, int lane) {
        float hid[HPL];
        {
            float e[HPL], d[HPL];
#pragma unroll
            for (int k = 0; k < HPL; ++k) e[k] = ex2_ftz(zs[k]);
            vadd_s<HPL>(d, e, 1.0f);
#pragma unroll
            for (int k = 0; k < HPL; ++k) hid[k] = rcp_ftz(d[k]);                 // R:52-53
        }
        // ---- output layer (chain)
        float t[O];
        float g_all[HG];
        if constexpr (GATHER) {
#pragma unroll
            for (int h = 0; h < H; ++h) g_all[h] = __shfl_sync(0xffffffffu, hid[0], h);
#pragma unroll
            for (int o = 0; o < O; ++o) {
                float a0 = b2l[o], a1 = 0.0f;                // two partial sums: shorter dependent chain
#pragma unroll
                for (int h = 0; h + 1 < H; h += 2) {
                    a0 = fmaf(g_all[h], -kL2E * w2f[h][o], a0);
                    a1 = fmaf(g_all[h + 1], -kL2E * w2f[h + 1][o], a1);
                }
                if (H & 1) a0 = fmaf(g_all[H - 1], -kL2E * w2f[H - 1][o], a0);
                t[o] = a0 + a1;
            }
        } else {
            if constexpr (FIX) {
                int sraw[O];
#pragma unroll
                for (int o = 0; o < O; ++o) {
                    float a0 = kMagic;                       // every partial lands on the integer grid
#pragma unroll
                    for (int k = 0; k < HPL; ++k) a0 = fmaf(hid[k], w2q[o][k], a0);
                    sraw[o] = __reduce_add_sync(0xffffffffu, __float_as_int(a0));    // wraps mod 2^32
                }
#pragma unroll
                for (int o = 0; o < O; ++o) t[o] = fmaf((float)(int)((unsigned int)sraw[o] - (unsigned int)kMagicSum32), cdec, b2l[o]);
            } else {
#pragma unroll
                for (int o = 0; o < O; ++o) {
                    float a0 = hid[0] * w2[o][0];
#pragma unroll
                    for (int k = 1; k < HPL; ++k) a0 = fmaf(hid[k], w2[o][k], a0);
                    t[o] = fmaf(warp_sum_shfl(a0, 5), -kL2E, b2l[o]);
                }
            }
        }
        // ---- independent of the chain: stale scaled pre-activation of the next row, xn.x + 1,
        //      hid (1 - hid) W2 with the PRE-update W2 (R:59)
        float zn[HPL];
        preact(xn, zn);
        float c = 1.0f;
#pragma unroll
        for (int i = 0; i < I; ++i) c = fmaf(xn[i], x[i], c);
        const float ccl = (-kL2E * lr) * c;
        float g[HPL], nh[HPL];
        vmul_s<HPL>(nh, hid, -1.0f);
        vfma<HPL>(g, nh, hid, hid);                                               // hid - hid^2
        float gw[O][HPL], ks[O][HPL];
#pragma unroll
        for (int o = 0; o < O; ++o) {
            vmul<HPL>(gw[o], g, w2[o]);
            vmul_s<HPL>(ks[o], gw[o], ccl);
        }
        // ---- output sigmoid and delta (chain)
        float od[O];
#pragma unroll
        for (int o = 0; o < O; ++o) {
            const float out = rcp_ftz(1.0f + ex2_ftz(t[o]));                       // R:54-55
            float dd;
            if constexpr (TASK == kTaskCls) dd = ((int)yv == o) ? 1.0f : 0.0f;      // C:73-75 one-hot
            else dd = yv;                                                          // O == 1 (R:132)
            od[o] = (dd - out) * fmaf(-out, out, out);                             // R:58
        }
        // ---- next row's scaled pre-activation (chain: O FFMAs)
#pragma unroll
        for (int o = 0; o < O; ++o) vfma_s<HPL>(zn, ks[o], od[o], zn);
#pragma unroll
        for (int k = 0; k < HPL; ++k) zs[k] = zn[k];
        // ---- updates (off the chain)
        float lh[HPL];
        vmul_s<HPL>(lh, gw[0], od[0]);
#pragma unroll
        for (int o = 1; o < O; ++o) vfma_s<HPL>(lh, gw[o], od[o], lh);
        vmul_s<HPL>(lh, lh, lr);                                                   // lr * hid_delta (R:59)
#pragma unroll
        for (int i = 0; i < I; ++i) vfma_s<HPL>(w1[i], lh, x[i], w1[i]);           // R:74-76
        vfma_s<HPL>(b1, lh, -1.0f, b1);                                            // R:77-78
#pragma unroll
        for (int o = 0; o < O; ++o) {
            const float lo = lr * od[o];
            vfma_s<HPL>(w2[o], hid, lo, w2[o]);                                    // R:67-69
            b2[o] -= lo;                                                           // R:70-71
            b2l[o] = kL2E * b2[o];
            if constexpr (GATHER) {
#pragma unroll
                for (int h = 0; h < H; ++h) w2f[h][o] = fmaf(lo, g_all[h], w2f[h][o]);
            }
        }
        refresh(lane);
    }
};

// shared-memory row fetch (the SGD warp always reads rows from shared memory: the staged copy or a TMA tile)
template <int IP>
__device__ __forceinline__ void lds_row(uint32_t xaddr, float (&x)[IP]) {
#pragma unroll
    for (int q = 0; q < IP / 4; ++q)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(x[4 * q]), "=f"(x[4 * q + 1]), "=f"(x[4 * q + 2]), "=f"(x[4 * q + 3])
                     : "r"(xaddr + 16u * q));
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// Streaming state of the SGD warp when the training set does not fit in shared memory:
// two TMA tiles, one mbarrier each.
struct SgdStream {
    float *tile_x0, *tile_x1, *tile_y0, *tile_y1;
    uint64_t *bar0, *bar1;
    uint32_t parity0, parity1;
};

// One epoch of online SGD over the training rows in order.  Called by the serial warp only.
//
// Row r needs row r+1 (look-ahead pre-activation) in registers when it starts, so row r+2 is
// fetched from shared memory at the top of row r: three register buffers rotate A -> B -> C and
// the hot loop is unrolled by three (no register moves, and the scheduler sees the loop-carried
// chain across rows).  Streamed training sets (two 128-row TMA tiles) cut the epoch into chunks at
// the rows where something else has to happen: r = 128k (refill the tile buffer that was just
// vacated) and r = 128k + 126 (the fetch of row r+2 crosses into the next tile: wait for it).
template <int I, int H, int O, int TASK>
__device__ __forceinline__ void sgd_pass(const float *w_in, float *w_out, const DataView &d, bool staged,
                                         float lr, SgdStream &st) {
    constexpr int IP = IPad<I>::value;
    constexpr uint32_t RB = IP * 4u;
    const int lane = threadIdx.x & 31;
    // opaque copies: kernel parameters would otherwise be re-read from the constant bank inside the
    // row loop (an LDC and its latency in series with the dependent chain)
    int n = d.n;
    asm volatile("" : "+r"(n));
    asm volatile("" : "+f"(lr));
    SgdWarp<I, H, O, TASK> net;
    net.load(w_in, lane);
    const int ntiles = (n + kTileRows - 1) / kTileRows;
    const uint32_t sx = smem_u32(d.x), sy = smem_u32(d.y);                       // staged copy
    const uint32_t tx0 = smem_u32(st.tile_x0), tx1 = smem_u32(st.tile_x1);       // streamed tiles
    const uint32_t ty0 = smem_u32(st.tile_y0), ty1 = smem_u32(st.tile_y1);
    auto xaddr = [&](int r) -> uint32_t {
        return staged ? sx + (uint32_t)r * RB : (((r / kTileRows) & 1) ? tx1 : tx0) + (uint32_t)(r % kTileRows) * RB;
    };
    auto yaddr = [&](int r) -> uint32_t {
        return staged ? sy + (uint32_t)r * 4u : (((r / kTileRows) & 1) ? ty1 : ty0) + (uint32_t)(r % kTileRows) * 4u;
    };
    auto issue = [&](int t) {
        const int rows = min(kTileRows, n - t * kTileRows);
        const uint32_t bx = (uint32_t)rows * RB;
        const uint32_t by = (uint32_t)((rows + 3) & ~3) * 4u;   // y is allocated padded to 4 floats
        if (lane == 0) {
            uint64_t *bar = (t & 1) ? st.bar1 : st.bar0;
            mbar_arrive_expect_tx(bar, bx + by);
            tma_load_1d((t & 1) ? st.tile_x1 : st.tile_x0, d.x + (size_t)t * kTileRows * IP, bx, bar);
            tma_load_1d((t & 1) ? st.tile_y1 : st.tile_y0, d.y + (size_t)t * kTileRows, by, bar);
        }
    };
    auto wait = [&](int t) {
        if (t & 1) { mbar_wait(st.bar1, st.parity1); st.parity1 ^= 1u; }
        else { mbar_wait(st.bar0, st.parity0); st.parity0 ^= 1u; }
    };
    if (!staged) { issue(0); wait(0); }
    float A[IP], B[IP], C[IP], ya, yb, yc;
    float zs[SgdWarp<I, H, O, TASK>::HPL];
    lds_row<IP>(xaddr(0), A);
    ya = lds_f32(yaddr(0));
    lds_row<IP>(xaddr(min(1, n - 1)), B);
    yb = lds_f32(yaddr(min(1, n - 1)));
    net.preact(A, zs);
    int r = 0;
    while (r < n) {
        int e = n;
        if (!staged) {
            const int k = r % kTileRows;
            if (k == 0) {                        // every lane is past the previous tile: refill its buffer
                __syncwarp();
                if (r / kTileRows + 1 < ntiles) issue(r / kTileRows + 1);
            }
            if (k == kTileRows - 2 && r + 2 < n) wait((r + 2) / kTileRows);
            e = min(n, r - k + (k < kTileRows - 2 ? kTileRows - 2 : kTileRows));
        }
        if (!SgdWarp<I, H, O, TASK>::GATHER) e = min(e, r + kGuardRows);   // rows covered by one fixed-point scale
        int cnt = e - r;
        uint32_t px = xaddr(min(r + 2, n - 1)), py = yaddr(min(r + 2, n - 1));
        // rows r, r+1, r+2 per iteration; fetches r+2, r+3, r+4 (same tile by construction).  The scale
        // of the fixed-point sum is chosen once per chunk (its CREDUX would otherwise sit in the loop
        // header, serial with the chain), with the margin for kGuardRows rows.
        int iters = min(cnt, max(n - 4 - r + 2, 0)) / 3;   // every fetch stays below n
        const bool ok = net.set_scale(kGuardRows, lr);
        r += 3 * iters; cnt -= 3 * iters;
        if (ok) {
            for (; iters > 0; --iters) {
                lds_row<IP>(px, C); yc = lds_f32(py);
                net.template row<true>(A, ya, B, zs, lr, lane);
                lds_row<IP>(px + RB, A); ya = lds_f32(py + 4u);
                net.template row<true>(B, yb, C, zs, lr, lane);
                lds_row<IP>(px + 2u * RB, B); yb = lds_f32(py + 8u);
                net.template row<true>(C, yc, A, zs, lr, lane);
                px += 3u * RB; py += 12u;
            }
        } else {
            for (; iters > 0; --iters) {
                lds_row<IP>(px, C); yc = lds_f32(py);
                net.template row<false>(A, ya, B, zs, lr, lane);
                lds_row<IP>(px + RB, A); ya = lds_f32(py + 4u);
                net.template row<false>(B, yb, C, zs, lr, lane);
                lds_row<IP>(px + 2u * RB, B); yb = lds_f32(py + 8u);
                net.template row<false>(C, yc, A, zs, lr, lane);
                px += 3u * RB; py += 12u;
            }
        }
        while (cnt > 0) {
            const int q = min(r + 2, n - 1);     // past the end: any finite row (the look-ahead result is unused)
            lds_row<IP>(xaddr(q), C); yc = lds_f32(yaddr(q));
            if (ok) net.template row<true>(A, ya, B, zs, lr, lane);
            else net.template row<false>(A, ya, B, zs, lr, lane);
#pragma unroll
            for (int i = 0; i < IP; ++i) { A[i] = B[i]; B[i] = C[i]; }
            ya = yb; yb = yc;
            ++r; --cnt;
        }
    }
    net.store(w_out, lane);
}

// ==========================================================================================
// K2 (wide hidden layers): the serial recurrence run by a TEAM of NTT threads (H > 64).
//
// One warp cannot hold a 256-wide layer (8 hidden units x (I + O + 1) weights per lane spill), so the
// recurrence of SgdWarp is spread over NTT = 128 threads, UPT = ceil(H / NTT) hidden units per thread, ROW
// VIEW ONLY: thread t keeps the W1 columns, B1 and the W2 rows of its units in registers.  Per row:
//
//     hid = sigmoid(zs)                                   per unit               (EX2, FADD, RCP)
//     partial[o] = sum over the thread's units hid W2q    packed over o pairs    (FFMA2 on the magic constant)
//     REDUX.SUM per output                                O per warp, pipelined  (block fixed point, as SgdWarp)
//     lane 0 -> s_part[row parity][warp][o];  ONE team barrier
//     lane o of EVERY warp: sum of the NWT warp sums (integer: exact, order free) -> out, out_delta, lr out_delta,
//              B2 -- redundantly per warp, so the broadcast of lr out_delta[0..O) to the lanes is warp-local
//              (STS, __syncwarp, LDS.128) and needs no second team barrier
//     lr hid_delta = hid (1 - hid) sum_o lr out_delta[o] W2[.][o]   with the pre-update W2 (R:59)
//     zs(next row) = stale pre-activation + lr hid_delta * -log2e (x_next . x + 1)       (one FFMA on the chain)
//     W2 += hid (x) lr out_delta (needed by the next row's partial sums); W1 / B1 of THIS row are applied in the
//     shadow of the NEXT row's REDUX latency, followed by the stale pre-activation of the row after it.
//
// Round 1's team kept W2 in a second, column-major view (one warp per output) and needed two barriers per row:
// 138 instructions per thread and row on 256 threads, 1137 cycles per row with two temperatures per SM.  Here:
// one view, one barrier, ~125 instructions per thread and row on 128 threads.
// ==========================================================================================
template <int H>
struct UseSgdTeam {
    static constexpr bool value = H > 64;
};
constexpr int kTeamThreads = 128;
__host__ __device__ constexpr int team_smem_floats(int H, int O) {
    // s_part [2][8 warps][OP] (int), s_od [8][OP], s_max [8], s_c [kTileRows]
    return 3 * 8 * ((O + 3) & ~3) + 8 + kTileRows + 4;
}

template <int NTT>
__device__ __forceinline__ void team_bar() {
    asm volatile("bar.sync 1, %0;" ::"n"(NTT) : "memory");
}

template <int I, int H, int O, int TASK, int NTT>
__device__ __forceinline__ void sgd_pass_team(const float *w_in, float *w_out, const DataView &d, bool staged, float lr,
                                              SgdStream &st, float *s_team) {
    constexpr int IP = IPad<I>::value;
    constexpr uint32_t RBY = IP * 4u;
    constexpr int NWT = NTT / 32;
    constexpr int UPT = (H + NTT - 1) / NTT;      // hidden units per thread: h = tid + NTT * k
    constexpr int OP = (O + 3) & ~3;              // outputs padded to whole LDS.128 / f32x2 pairs
    constexpr int OJ = OP / 2;
    constexpr int oW2 = I * H, oB1 = I * H + H * O, oB2 = I * H + H * O + H;
    static_assert(NTT % 32 == 0 && NWT <= 8 && O <= 32, "team geometry");
    int *s_part = reinterpret_cast<int *>(s_team);            // [2][NWT][OP]
    float *s_od = s_team + 2 * 8 * OP;                        // [NWT][OP]  lr * out_delta of the row, per warp
    unsigned int *s_max = reinterpret_cast<unsigned int *>(s_od + 8 * OP);   // [NWT]
    float *s_c = reinterpret_cast<float *>(s_max + 8);        // [kTileRows]  x_{r+1}.x_r + 1 of the resident tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n = d.n;
    asm volatile("" : "+r"(n));                   // keep loop invariants in registers (no LDC in the row loop)
    asm volatile("" : "+f"(lr));

    // ---- weights of this thread's units (layout a1, R:80-90)
    float w1[I][UPT], b1[UPT];
    // W2 rows, outputs in pairs, kept ONLY in block-fixed-point scale: w2s = fscale * W2.  fscale is a power of two, so
    // every product, sum and update below is the unscaled one times fscale exactly (scaling by a power of two commutes
    // with fp32 rounding), and W2 = w2s / fscale at the end is bit for bit what an unscaled copy would hold.
    f2_t w2s[UPT][OJ];
    float b2 = 0.0f, b2l = 0.0f;                  // lane o of every warp: B2[o] (identical updates in every warp)
#pragma unroll
    for (int k = 0; k < UPT; ++k) {
        const int h = tid + NTT * k;
        const bool act = h < H;
#pragma unroll
        for (int i = 0; i < I; ++i) w1[i][k] = act ? w_in[i * H + h] : 0.0f;
        // no hidden unit: z = x.0 - 1e4 -> sigmoid = 0 exactly, so every update of the slot is an exact zero
        b1[k] = act ? w_in[oB1 + h] : 1.0e4f;
#pragma unroll
        for (int j = 0; j < OJ; ++j) {
            const float a = (act && 2 * j < O) ? w_in[oW2 + h * O + 2 * j] : 0.0f;
            const float b = (act && 2 * j + 1 < O) ? w_in[oW2 + h * O + 2 * j + 1] : 0.0f;
            w2s[k][j] = pack2(a, b);
        }
    }
    if (lane < O) { b2 = w_in[oB2 + lane]; b2l = kL2E * b2; }
    team_bar<NTT>();   // w_in may alias w_out; everybody has read its share

    // ---- data addressing: staged copy or two TMA tiles
    const int ntiles = (n + kTileRows - 1) / kTileRows;
    const uint32_t sx = smem_u32(d.x), sy = smem_u32(d.y);
    const uint32_t tx0 = smem_u32(st.tile_x0), tx1 = smem_u32(st.tile_x1);
    const uint32_t ty0 = smem_u32(st.tile_y0), ty1 = smem_u32(st.tile_y1);
    auto xaddr = [&](int r) -> uint32_t {
        return staged ? sx + (uint32_t)r * RBY : (((r / kTileRows) & 1) ? tx1 : tx0) + (uint32_t)(r % kTileRows) * RBY;
    };
    auto yaddr = [&](int r) -> uint32_t {
        return staged ? sy + (uint32_t)r * 4u : (((r / kTileRows) & 1) ? ty1 : ty0) + (uint32_t)(r % kTileRows) * 4u;
    };
    auto issue = [&](int t) {
        const int rows = min(kTileRows, n - t * kTileRows);
        const uint32_t bx = (uint32_t)rows * RBY;
        const uint32_t by = (uint32_t)((rows + 3) & ~3) * 4u;
        if (tid == 0) {
            uint64_t *bar = (t & 1) ? st.bar1 : st.bar0;
            mbar_arrive_expect_tx(bar, bx + by);
            tma_load_1d((t & 1) ? st.tile_x1 : st.tile_x0, d.x + (size_t)t * kTileRows * IP, bx, bar);
            tma_load_1d((t & 1) ? st.tile_y1 : st.tile_y0, d.y + (size_t)t * kTileRows, by, bar);
        }
    };
    auto wait = [&](int t) {
        if (t & 1) { mbar_wait(st.bar1, st.parity1); st.parity1 ^= 1u; }
        else { mbar_wait(st.bar0, st.parity0); st.parity0 ^= 1u; }
    };

    float zs[UPT];            // ex2-domain pre-activations of this thread's units for the current row
    float fscale = 1.0f, inv_fscale = 1.0f, cdec = -kL2E;
    // Fixed-point scale of the output sums for the next `rows` rows (SgdWarp::set_scale, here agreed by the whole
    // team: one extra barrier per kGuardRows rows).  |partial[o]| of a thread <= sum_k |W2[k][o]| because hid is in
    // [0,1]; one row moves each |W2| entry by at most lr/4.  2^E > bound -> scale 2^(22-E): a thread's scaled partial
    // stays below 2^22 (integer-exact on the magic constant) and 32 lanes x NWT warps of it cannot overflow int32.
    // Non-finite weights poison the decode factor so that NaN propagates as in the reference.
    auto set_scale = [&](int rows) {
        float m = 0.0f;
#pragma unroll
        for (int j = 0; j < OJ; ++j) {
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int k = 0; k < UPT; ++k) {
                float x0, x1;
                unpack2(w2s[k][j], x0, x1);
                a0 += fabsf(x0); a1 += fabsf(x1);
            }
            a0 *= inv_fscale; a1 *= inv_fscale;                     // back to the scale of W2 (exact)
            m = fmaxf(m, fmaxf(a0, a1));
            m = (a0 != a0 || a1 != a1) ? __int_as_float(0x7fc00000) : m;
        }
        m = fmaxf(m + (float)(rows * UPT) * 0.25f * fabsf(lr), 9.765625e-4f);
        const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(m));     // NaN / Inf bits compare high
        if (lane == 0) s_max[warp] = wm;
        team_bar<NTT>();
        unsigned int mb = 0u;
#pragma unroll
        for (int w = 0; w < NWT; ++w) mb = max(mb, s_max[w]);
        const unsigned int ef = mb >> 23;
        const bool ok = ef < 200u;
        const unsigned int sf = ok ? 275u - ef : 127u;
        const float fnew = __uint_as_float(sf << 23);
        const float ratio = fnew * inv_fscale;                     // a power of two
        fscale = fnew;
        inv_fscale = __uint_as_float((254u - sf) << 23);
        cdec = ok ? -kL2E * inv_fscale : __int_as_float(0x7fc00000);
        const f2_t r2 = pack2(ratio, ratio);
#pragma unroll
        for (int k = 0; k < UPT; ++k)
#pragma unroll
            for (int j = 0; j < OJ; ++j) w2s[k][j] = mul2(w2s[k][j], r2);
    };

    // carried from one row to the next: the W1 / B1 update of the previous row is applied in the next row
    float lh_p[UPT];
#pragma unroll
    for (int k = 0; k < UPT; ++k) lh_p[k] = 0.0f;
    int par = 0;              // row parity: s_part is double buffered

    // One row.  The current row enters through zs (its pre-activations) and y.  Input rows are NOT kept in registers
    // from one row to the next: the previous row (its W1 update is still pending) and the look-ahead row are read from
    // shared memory where they are used (broadcast LDS.128: the port is idle otherwise) -- 32 registers that the
    // 168-register budget of the five-warp tcgen05 geometry does not have.  xprev = shared-memory address of the previous
    // row; the pending update is flushed before a streamed tile buffer is recycled.
    uint32_t xprev = 0u;
    auto row = [&](const bool boundary, uint32_t xcur_addr, uint32_t y_addr, uint32_t xn_addr, float c_in, bool has_next) {
        const float yv = lds_f32(y_addr);
        // ---- hidden activations (R:52-53)
        float hid[UPT];
#pragma unroll
        for (int k = 0; k < UPT; ++k) hid[k] = rcp_ftz(1.0f + ex2_ftz(zs[k]));
        // ---- partial output sums on the magic constant, outputs in pairs; integer REDUX per output
        int sraw[OP];
#pragma unroll
        for (int j = 0; j < OJ; ++j) {
            f2_t a = pack2(kMagic, kMagic);
#pragma unroll
            for (int k = 0; k < UPT; ++k) a = fma2(pack2(hid[k], hid[k]), w2s[k][j], a);
            float a0, a1;
            unpack2(a, a0, a1);
            sraw[2 * j] = (2 * j < O) ? __reduce_add_sync(0xffffffffu, __float_as_int(a0)) : 0;
            sraw[2 * j + 1] = (2 * j + 1 < O) ? __reduce_add_sync(0xffffffffu, __float_as_int(a1)) : 0;
        }
        // ---- in the shadow of the REDUX latency: W1 / B1 update of the PREVIOUS row (R:74-78) ...
        {
            float xp[IP];
            lds_row<IP>(xprev, xp);
#pragma unroll
            for (int i = 0; i < I; ++i) vfma_s<UPT>(w1[i], lh_p, xp[i], w1[i]);
        }
#pragma unroll
        for (int k = 0; k < UPT; ++k) b1[k] -= lh_p[k];
        // ... then the look-ahead row takes its place: stale pre-activation of the next row, with the weights
        // BEFORE this row's update (the previous row's has just been applied)
        float cr = c_in;
        float xn[IP];
        if (boundary) {                               // literal at every call site: folded after inlining
            if (has_next) {
                float xc[IP];
                lds_row<IP>(xn_addr, xn);
                lds_row<IP>(xcur_addr, xc);
                cr = 1.0f;
#pragma unroll
                for (int i = 0; i < I; ++i) cr = fmaf(xc[i], xn[i], cr);
            } else {
#pragma unroll
                for (int i = 0; i < IP; ++i) xn[i] = 0.0f;
                cr = 1.0f;
            }
        } else {
            lds_row<IP>(xn_addr, xn);
        }
        float zn[UPT];
        {
            float t[UPT];
            vmul_s<UPT>(t, b1, -1.0f);
#pragma unroll
            for (int i = 0; i < I; ++i) vfma_s<UPT>(t, w1[i], xn[i], t);
            vmul_s<UPT>(zn, t, -kL2E);
        }
        const float ccl = -kL2E * cr;
        float g[UPT];
#pragma unroll
        for (int k = 0; k < UPT; ++k) g[k] = fmaf(-hid[k], hid[k], hid[k]) * inv_fscale;   // hid (1 - hid), and the un-scaling of the W2 dot product below
        // ---- publish this warp's sums; the ONE team barrier of the row
        if (lane == 0) {
            int4 *dst = reinterpret_cast<int4 *>(s_part + (par * 8 + warp) * OP);
#pragma unroll
            for (int q = 0; q < OP / 4; ++q) dst[q] = make_int4(sraw[4 * q], sraw[4 * q + 1], sraw[4 * q + 2], sraw[4 * q + 3]);
        }
        team_bar<NTT>();
        // ---- lane o: output o (every warp computes all of them: the broadcast below stays inside the warp)
        {
            const int o = lane < OP ? lane : OP - 1;
            unsigned int us = 0u;
#pragma unroll
            for (int w = 0; w < NWT; ++w) us += (unsigned int)s_part[(par * 8 + w) * OP + o];
            const float t = fmaf((float)(int)(us - (unsigned int)NWT * (unsigned int)kMagicSum32), cdec, b2l);
            const float out = rcp_ftz(1.0f + ex2_ftz(t));                          // R:54-55
            float dd;
            if constexpr (TASK == kTaskCls) dd = ((int)yv == lane) ? 1.0f : 0.0f;  // C:73-75 one-hot
            else dd = yv;                                                          // O == 1 (R:132)
            const float lo = (lane < O) ? lr * ((dd - out) * fmaf(-out, out, out)) : 0.0f;   // lr * out_delta (R:58)
            b2 -= lo;                                                              // R:70-71
            b2l = kL2E * b2;
            if (lane < OP) s_od[warp * OP + lane] = lo;
        }
        __syncwarp();
        f2_t lo2[OJ];
#pragma unroll
        for (int q = 0; q < OP / 4; ++q) {
            const float4 v = reinterpret_cast<const float4 *>(s_od + warp * OP)[q];
            lo2[2 * q] = pack2(v.x, v.y); lo2[2 * q + 1] = pack2(v.z, v.w);
        }
        // ---- lr * hid_delta with the PRE-update W2 (R:59); next row's pre-activation (chain: one FFMA)
        float lh[UPT];
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
            f2_t a = mul2(w2s[k][0], lo2[0]);
#pragma unroll
            for (int j = 1; j < OJ; ++j) a = fma2(w2s[k][j], lo2[j], a);
            float a0, a1;
            unpack2(a, a0, a1);
            lh[k] = (a0 + a1) * g[k];
            zs[k] = fmaf(lh[k], ccl, zn[k]);
            lh_p[k] = lh[k];
        }
        // ---- W2 += hid (x) lr out_delta (R:67-69), in the scaled domain: the next row's partial sums need it
        const f2_t fs2 = pack2(fscale, fscale);
#pragma unroll
        for (int j = 0; j < OJ; ++j) lo2[j] = mul2(lo2[j], fs2);
#pragma unroll
        for (int k = 0; k < UPT; ++k) {
            const f2_t h2 = pack2(hid[k], hid[k]);
#pragma unroll
            for (int j = 0; j < OJ; ++j) w2s[k][j] = fma2(h2, lo2[j], w2s[k][j]);
        }
        par ^= 1;
        xprev = xcur_addr;
    };
    // applies the W1 / B1 update still pending after the last row processed (xprev)
    auto flush = [&]() {
        float xp[IP];
        lds_row<IP>(xprev, xp);
#pragma unroll
        for (int i = 0; i < I; ++i) vfma_s<UPT>(w1[i], lh_p, xp[i], w1[i]);
#pragma unroll
        for (int k = 0; k < UPT; ++k) { b1[k] -= lh_p[k]; lh_p[k] = 0.0f; }
    };

    // ---- rows
    constexpr bool PlainRow = false, TileEnd = true;
    if (!staged) { issue(0); wait(0); }
    xprev = xaddr(0);                              // "previous row" of row 0: nothing pending (lh_p = 0), any finite row
    {
        float x0[IP], t[UPT];
        lds_row<IP>(xaddr(0), x0);
        vmul_s<UPT>(t, b1, -1.0f);
#pragma unroll
        for (int i = 0; i < I; ++i) vfma_s<UPT>(t, w1[i], x0[i], t);
        vmul_s<UPT>(zs, t, -kL2E);
    }
    int r = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int base = t * kTileRows;
        const int rows = min(kTileRows, n - base);
        // tile t is resident.  x_{r+1}.x_r + 1 for its rows (the last one needs the next tile: done in its row)
        for (int q = tid; q < rows - 1; q += NTT) {
            float u[IP], v[IP];
            lds_row<IP>(xaddr(base + q), u);
            lds_row<IP>(xaddr(base + q + 1), v);
            float c = 1.0f;
#pragma unroll
            for (int i = 0; i < I; ++i) c = fmaf(u[i], v[i], c);
            s_c[q] = c;
        }
        team_bar<NTT>();          // s_c visible; every thread is past its reads of tile t-1
        if (!staged && t + 1 < ntiles) issue(t + 1);
        const int last = base + rows - 1;            // the tile's last row looks ahead into the next tile
        while (last - r >= 2) {                      // pairs of rows whose look-ahead rows are in this tile
            int pairs = min(kGuardRows, last - r) / 2;
            set_scale(kGuardRows);
            uint32_t px = xaddr(r + 1), py = yaddr(r);
            const float *pc = s_c + (r - base);
            r += 2 * pairs;
            for (; pairs > 0; --pairs) {
                row(PlainRow, px - RBY, py, px, pc[0], true);             // current row r, look-ahead r + 1
                row(PlainRow, px, py + 4u, px + RBY, pc[1], true);        // current row r + 1, look-ahead r + 2
                px += 2u * RBY; py += 8u; pc += 2;
            }
        }
        set_scale(2);
        const bool has_next = last + 1 < n;
        if (has_next && !staged) wait(t + 1);        // issued 127 rows ago
        const uint32_t xnext = xaddr(has_next ? last + 1 : last);
        if (last - r == 1) row(PlainRow, xaddr(r), yaddr(r), xaddr(r + 1), s_c[r - base], true);
        row(TileEnd, xaddr(last), yaddr(last), xnext, 1.0f, has_next);
        r = last + 1;
        flush();                                     // before the next tile's refill recycles this tile's buffer
    }

#pragma unroll
    for (int k = 0; k < UPT; ++k) {
        const int h = tid + NTT * k;
        if (h < H) {
#pragma unroll
            for (int i = 0; i < I; ++i) w_out[i * H + h] = w1[i][k];
#pragma unroll
            for (int j = 0; j < OJ; ++j) {
                float a0, a1;
                unpack2(w2s[k][j], a0, a1);
                if (2 * j < O) w_out[oW2 + h * O + 2 * j] = a0 * inv_fscale;
                if (2 * j + 1 < O) w_out[oW2 + h * O + 2 * j + 1] = a1 * inv_fscale;
            }
            w_out[oB1 + h] = b1[k];
        }
    }
    if (warp == 0 && lane < O) w_out[oB2 + lane] = b2;
    team_bar<NTT>();
}

// ==========================================================================================
// K1: row-parallel forward + likelihood partial sums.  Thread t of a team of nt threads takes
// rows t, t+nt, ...  (RB rows at a time so each broadcast weight load is reused RB times).
//   regression:      s0 += (y - fx)^2
//   classification:  s0 += log softmax(out)[label] (C:215-219), s1 += (argmax - y)^2 (C:212),
//                    correct += (argmax == y) (C:200-207)
// ==========================================================================================
template <int I, int H, int O, int TASK, int RB, bool PRECISE, bool WRITE>
__device__ __forceinline__ void lik_rows_impl(const float *__restrict__ w, const DataView &d, int t, int nt,
                                              double &s0, double &s1, int &correct, float *fx_out,
                                              float *prob_out) {
    constexpr int IP = IPad<I>::value;
    constexpr int oW2 = I * H, oB1 = I * H + H * O, oB2 = I * H + H * O + H;
    constexpr int HU = (H <= 16) ? H : 4;
    for (int r0 = t * RB; r0 < d.n; r0 += nt * RB) {
        float x[RB][IP];
        float acc[RB][O];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = min(r0 + b, d.n - 1);
            const float4 *xr = reinterpret_cast<const float4 *>(d.x + (size_t)r * IP);
#pragma unroll
            for (int q = 0; q < IP / 4; ++q) {
                const float4 v = xr[q];
                x[b][4 * q] = v.x; x[b][4 * q + 1] = v.y; x[b][4 * q + 2] = v.z; x[b][4 * q + 3] = v.w;
            }
#pragma unroll
            for (int o = 0; o < O; ++o) acc[b][o] = -w[oB2 + o];
        }
#pragma unroll HU
        for (int h = 0; h < H; ++h) {
            float w1h[I], w2h[O];
#pragma unroll
            for (int i = 0; i < I; ++i) w1h[i] = w[i * H + h];
#pragma unroll
            for (int o = 0; o < O; ++o) w2h[o] = w[oW2 + h * O + o];
            const float nb = -w[oB1 + h];
            float zz[RB], hid[RB];
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                float z = nb;
#pragma unroll
                for (int i = 0; i < I; ++i) z = fmaf(x[b][i], w1h[i], z);
                zz[b] = z;
            }
            if constexpr (PRECISE) {
#pragma unroll
                for (int b = 0; b < RB; ++b) hid[b] = sigmoid_precise(zz[b]);
            } else {
                sigmoid_group<RB>(zz, hid);
            }
#pragma unroll
            for (int b = 0; b < RB; ++b) {
#pragma unroll
                for (int o = 0; o < O; ++o) acc[b][o] = fmaf(hid[b], w2h[o], acc[b][o]);
            }
        }
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = r0 + b;
            if (r < d.n) {
                const float yv = d.y[r];
                if constexpr (TASK == kTaskReg) {
                    const float fx = sigmoid_sel<PRECISE>(acc[b][0]);             // R:55, R:132
                    const float e = yv - fx;
                    s0 += (double)(e * e);
                    if constexpr (WRITE) fx_out[r] = fx;
                } else {
                    float out[O];
                    int am = 0;
                    float se = 0.0f;
#pragma unroll
                    for (int o = 0; o < O; ++o) {
                        out[o] = sigmoid_sel<PRECISE>(acc[b][o]);
                        se += expf(out[o]);            // C:108-110
                    }
#pragma unroll
                    for (int o = 1; o < O; ++o) am = (out[o] > out[am]) ? o : am; // np.argmax: first max
                    const int lab = (int)yv;
                    float ol = out[0];
#pragma unroll
                    for (int o = 1; o < O; ++o) ol = (o == lab) ? out[o] : ol;
                    s0 += (double)(ol - logf(se));
                    const float e = (float)am - yv;
                    s1 += (double)(e * e);
                    correct += ((float)am == yv) ? 1 : 0;
                    if constexpr (WRITE) {
                        fx_out[r] = (float)am;
                        if (prob_out) {
#pragma unroll
                            for (int o = 0; o < O; ++o) prob_out[(size_t)r * O + o] = expf(out[o]) / se;
                        }
                    }
                }
            }
        }
    }
}

template <int I, int H, int O, int TASK, bool PRECISE, bool WRITE>
__device__ __forceinline__ void lik_rows(const float *__restrict__ w, const DataView &d, int t, int nt, double &s0,
                                         double &s1, int &correct, float *fx_out = nullptr,
                                         float *prob_out = nullptr) {
    constexpr int IP = IPad<I>::value;
    // row blocking only pays when every thread has several rows and the registers allow it
    constexpr int RBMAX = (IP * 4 + O * 4 <= 96) ? 4 : ((IP * 2 + O * 2 <= 96) ? 2 : 1);
    if (RBMAX >= 4 && d.n >= 8 * nt) lik_rows_impl<I, H, O, TASK, RBMAX, PRECISE, WRITE>(w, d, t, nt, s0, s1, correct, fx_out, prob_out);
    else if (RBMAX >= 2 && d.n >= 4 * nt) lik_rows_impl<I, H, O, TASK, 2, PRECISE, WRITE>(w, d, t, nt, s0, s1, correct, fx_out, prob_out);
    else lik_rows_impl<I, H, O, TASK, 1, PRECISE, WRITE>(w, d, t, nt, s0, s1, correct, fx_out, prob_out);
}

// ==========================================================================================
// K1 (chain kernel): the same pass with the instruction count cut down.  ncu on the random-walk
// step of 4-64-1 showed 12.5 issued instructions per hidden-unit sigmoid with the issue slots 64 %
// busy, so the pass is rewritten around a per-proposal "likelihood layout" of the weights in shared
// memory, one 16-byte-aligned record per hidden unit:
//     [ -log2(e) W1[0..I)[h],  log2(e) B1[h],  -log2(e) W2[h][0..O),  pad ]
// (one or two LDS.128 per hidden unit instead of I + O + 1 scalar loads, no scaling multiplies in
// the loop, pre-activations born in the ex2 domain), rows processed in PAIRS through the packed
// FFMA2 / FMUL2 / FADD2 instructions, and one MUFU.RCP per four sigmoids (sigmoid_group).  For
// 4-64-1 that is 29 instructions and 5 MUFU per 4 rows x 1 hidden unit (was 50 and 5).
// ==========================================================================================
template <int I, int O>
struct LikLayout {
    static constexpr int LW = (I + 1 + O + 3) & ~3;
};

// element j of the weight vector (layout a1) with value v -> its slot in the likelihood layout
template <int I, int H, int O>
__device__ __forceinline__ void lik_scatter(float *lw, int j, float v) {
    constexpr int LW = LikLayout<I, O>::LW;
    constexpr int oW2 = I * H, oB1 = I * H + H * O, oB2 = I * H + H * O + H;
    if (j < oW2) { const int i = j / H, h = j - i * H; lw[h * LW + i] = -kL2E * v; }
    else if (j < oB1) { const int q = j - oW2, h = q / O, o = q - h * O; lw[h * LW + I + 1 + o] = -kL2E * v; }
    else if (j < oB2) { lw[(j - oB1) * LW + I] = kL2E * v; }
}
template <int I, int H, int O>
__device__ __forceinline__ void lik_prepare(float *lw, const float *w, int t, int nt) {
    for (int j = t; j < I * H + H * O + H; j += nt) lik_scatter<I, H, O>(lw, j, w[j]);
}

// Sigmoids of NP row pairs from their ex2-domain pre-activations (sigmoid(z), zs = -log2(e) z):
// one MUFU.EX2 each, ONE MUFU.RCP per four (per two when NP == 1).
template <int NP>
__device__ __forceinline__ void sigmoid_pairs(const f2_t (&zs)[NP], f2_t (&s)[NP]) {
    constexpr float kClamp = 30.0f;     // the product of up to four denominators must not overflow
    f2_t d[NP];
    const f2_t one = pack2(1.0f, 1.0f);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        float a, b;
        unpack2(zs[j], a, b);
        d[j] = add2(pack2(ex2_ftz(fminf(a, kClamp)), ex2_ftz(fminf(b, kClamp))), one);
    }
#pragma unroll
    for (int j = 0; j + 1 < NP; j += 2) {
        float pa, pb;
        unpack2(mul2(d[j], d[j + 1]), pa, pb);             // {d0 d2, d1 d3}
        const float r = rcp_ftz(pa * pb);
        const f2_t qs = pack2(r * pb, r * pa);
        s[j] = mul2(qs, d[j + 1]);                         // {1/d0, 1/d1}
        s[j + 1] = mul2(qs, d[j]);                         // {1/d2, 1/d3}
    }
    if (NP & 1) {
        float a, b;
        unpack2(d[NP - 1], a, b);
        const float r = rcp_ftz(a * b);
        s[NP - 1] = pack2(r * b, r * a);
    }
}

template <int I, int H, int O, int TASK, int RB>
__device__ __forceinline__ void lik_fast_impl(const float *__restrict__ lw, const float *__restrict__ w,
                                              const DataView &d, int t, int nt, double &s0, double &s1, int &correct) {
    constexpr int IP = IPad<I>::value;
    constexpr int LW = LikLayout<I, O>::LW;
    constexpr int oB2 = I * H + H * O + H;
    constexpr int HU = (H <= 16) ? H : 4;
    constexpr int NP = RB / 2;           // row pairs (RB == 1: scalar path)
    for (int r0 = t * RB; r0 < d.n; r0 += nt * RB) {
        float acc[O][RB];
        if constexpr (RB == 1) {
            float x[1][IP];
            {
                const float4 *xr = reinterpret_cast<const float4 *>(d.x + (size_t)r0 * IP);
#pragma unroll
                for (int q = 0; q < IP / 4; ++q) {
                    const float4 v = xr[q];
                    x[0][4 * q] = v.x; x[0][4 * q + 1] = v.y; x[0][4 * q + 2] = v.z; x[0][4 * q + 3] = v.w;
                }
            }
#pragma unroll
            for (int o = 0; o < O; ++o) acc[o][0] = kL2E * w[oB2 + o];
#pragma unroll HU
            for (int h = 0; h < H; ++h) {
                float wl[LW];
                const float4 *wr = reinterpret_cast<const float4 *>(lw + h * LW);
#pragma unroll
                for (int q = 0; q < LW / 4; ++q) {
                    const float4 v = wr[q];
                    wl[4 * q] = v.x; wl[4 * q + 1] = v.y; wl[4 * q + 2] = v.z; wl[4 * q + 3] = v.w;
                }
                float zs = wl[I];
#pragma unroll
                for (int i = 0; i < I; ++i) zs = fmaf(x[0][i], wl[i], zs);
                const float hid = rcp_ftz(1.0f + ex2_ftz(zs));
#pragma unroll
                for (int o = 0; o < O; ++o) acc[o][0] = fmaf(hid, wl[I + 1 + o], acc[o][0]);
            }
        } else {
            // scalar loads (once per row block) so that each pair is born in an aligned register pair
            f2_t x2[NP][I], a2[NP][O];
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const float *xa = d.x + (size_t)min(r0 + 2 * j, d.n - 1) * IP;
                const float *xb = d.x + (size_t)min(r0 + 2 * j + 1, d.n - 1) * IP;
#pragma unroll
                for (int i = 0; i < I; ++i) x2[j][i] = pack2(xa[i], xb[i]);
#pragma unroll
                for (int o = 0; o < O; ++o) { const float b2l = kL2E * w[oB2 + o]; a2[j][o] = pack2(b2l, b2l); }
            }
#pragma unroll HU
            for (int h = 0; h < H; ++h) {
                float wl[LW];
                const float4 *wr = reinterpret_cast<const float4 *>(lw + h * LW);
#pragma unroll
                for (int q = 0; q < LW / 4; ++q) {
                    const float4 v = wr[q];
                    wl[4 * q] = v.x; wl[4 * q + 1] = v.y; wl[4 * q + 2] = v.z; wl[4 * q + 3] = v.w;
                }
                f2_t zs[NP], hid[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    zs[j] = pack2(wl[I], wl[I]);
#pragma unroll
                    for (int i = 0; i < I; ++i) zs[j] = fma2(x2[j][i], pack2(wl[i], wl[i]), zs[j]);
                }
                sigmoid_pairs<NP>(zs, hid);
#pragma unroll
                for (int j = 0; j < NP; ++j)
#pragma unroll
                    for (int o = 0; o < O; ++o) a2[j][o] = fma2(hid[j], pack2(wl[I + 1 + o], wl[I + 1 + o]), a2[j][o]);
            }
#pragma unroll
            for (int j = 0; j < NP; ++j)
#pragma unroll
                for (int o = 0; o < O; ++o) unpack2(a2[j][o], acc[o][2 * j], acc[o][2 * j + 1]);
        }
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const int r = r0 + b;
            if (r < d.n) {
                const float yv = d.y[r];
                if constexpr (TASK == kTaskReg) {
                    const float fx = rcp_ftz(1.0f + ex2_ftz(acc[0][b]));          // R:55, R:132
                    const float e = yv - fx;
                    s0 += (double)(e * e);
                } else {
                    float out[O];
                    int am = 0;
                    float se = 0.0f;
#pragma unroll
                    for (int o = 0; o < O; ++o) {
                        out[o] = rcp_ftz(1.0f + ex2_ftz(acc[o][b]));
                        se += expf(out[o]);            // C:108-110
                    }
#pragma unroll
                    for (int o = 1; o < O; ++o) am = (out[o] > out[am]) ? o : am; // np.argmax: first max
                    const int lab = (int)yv;
                    float ol = out[0];
#pragma unroll
                    for (int o = 1; o < O; ++o) ol = (o == lab) ? out[o] : ol;
                    s0 += (double)(ol - logf(se));
                    const float e = (float)am - yv;
                    s1 += (double)(e * e);
                    correct += ((float)am == yv) ? 1 : 0;
                }
            }
        }
    }
}

template <int I, int H, int O, int TASK>
__device__ __forceinline__ void lik_fast(const float *__restrict__ lw, const float *__restrict__ w, const DataView &d,
                                         int t, int nt, double &s0, double &s1, int &correct) {
    // row blocking only pays when every thread has several rows and the registers allow it
    constexpr int RBMAX = (I * 4 + O * 4 <= 96) ? 4 : ((I * 2 + O * 2 <= 96) ? 2 : 1);
    if (RBMAX >= 4 && d.n >= 8 * nt) lik_fast_impl<I, H, O, TASK, RBMAX>(lw, w, d, t, nt, s0, s1, correct);
    else if (RBMAX >= 2 && d.n >= 4 * nt) lik_fast_impl<I, H, O, TASK, 2>(lw, w, d, t, nt, s0, s1, correct);
    else lik_fast_impl<I, H, O, TASK, 1>(lw, w, d, t, nt, s0, s1, correct);
}

}  // namespace ptfnn
#include "ptfnn_tc.cuh"
namespace ptfnn {

// K5 applies to the wide-hidden specialisations whose geometry matches the tcgen05 tile (M = 128 rows,
// H = 256 hidden units in two N = 128 halves, one epilogue warp per TMEM lane quadrant + the MMA warp)
template <int I, int H, int O, int NT>
struct UseTc {
    static constexpr bool value = (H == 256 && NT == tc::kThreads && O <= tc::kN2);
};

// ==========================================================================================
// K3 + K4: the persistent chain kernel
// ==========================================================================================
struct ChainParams {
    // ---- configuration
    int R, Rg, replica_offset;     // local replicas, whole ladder, ladder index of local replica 0
    int S, swap_interval, swap_rule;
    int use_lg, crn, memo, external_swap, debug;
    int step_begin, step_end;      // steps [begin, end) this launch
    int round_begin;               // swap rounds completed before this launch
    int final_round;               // 1: after the last step run the left-over coordinator round
    int replay, replay_n;          // draws come from HBM arrays covering steps [step_begin, step_begin+replay_n)
    int staged;                    // datasets fit in shared memory
    double l_prob, pt_samples, sigma_sq, nu1, nu2;
    float lr, step_w, step_eta;
    uint64_t seed;
    const double *temperature;     // [R]
    // ---- data
    DataView train, test;
    const float *a_train, *a_test; // K5: UMMA A tiles of the two data sets (tc::pack_a_kernel), else 0
    // ---- chain state (global, one row per local replica)
    float *w;                      // [R][P]
    double *eta, *tau, *lik, *prior;
    int *n_acc, *init_count;
    double *last4;                 // [R][4]  rmse_train, rmse_test, acc_train, acc_test carried rows (R:420-423)
    float *gd_cache;               // [R][P]  langevin_gradient(w) memo (wide nets: the working copy too)
    float *pgd_buf;                // [R][P]  langevin_gradient(w_prop) of wide nets (0 otherwise)
    float *prop_buf;               // [R][P]  the proposal of the tcgen05 topologies (their shared memory holds the MMA operands)
    int *gd_valid;                 // [R]
    // ---- traces
    float *pos_w;                  // [R][S][P]
    double *lik_prop, *rmse_tr, *rmse_te, *acc_tr, *acc_te;   // [R][S]
    int *accept_list;              // [R][S]
    double *dbg_prior, *dbg_diff, *dbg_mh;                    // [R][S] (debug)
    uint8_t *dbg_acc;
    // ---- replay draws
    const float *lx, *z, *z_eta, *u, *u_swap;
    // ---- swap round
    float *pub_rows;               // [2][R][P+2]   (w, eta) published at the hand-shake (R:430-431); eta = the fp64 bit pattern in two words
    double *pub_lhood;             // [2][Rg]       lhood field (R:430 / C:439)
    GridBarrier *barrier;
    long long *swap_counters;      // {num_swap, total_swap_proposals}  (R:680-688)
    uint8_t *swap_log;             // [rounds][Rg-1]
    int max_rounds;
    int *swap_src;                 // [R] origin slot of the vector that ends in each local slot (per round)
    int P;
    // ---- speculative windows for ladders that leave CTA slots free (spec_k > 1): spec_k CTAs per temperature
    //      share the next steps out among themselves (chain_body), each step assuming the earlier ones rejected
    int spec_k;
    int spec_plan;                 // 0 = choose per window; 1 = "apart", 2 = "riding" (measurement knob, chain_body)
    GridBarrier *spec_bar;         // [R]         barrier of the CTAs of one temperature
    unsigned int *spec_flag;       // [R][kSpecWords]: per step of the window, then per CTA of the group (chain_body)
    // ---- multi-GPU ladder through peer memory (n_ranks > 1): every rank's pub_lhood / pub_rows / peer_flags
    //      are mapped into this process (CUDA IPC); entry q of the tables points at rank q's buffer
    int n_ranks, rank;
    double *peer_lhood[kMaxPeers];         // rank q's pub_lhood  [2][Rg]
    const float *peer_rows[kMaxPeers];     // rank q's pub_rows   [2][R][P+2]
    unsigned int *peer_flags[kMaxPeers];   // rank q's arrival flags [kMaxPeers] + heartbeat; flags[j] = peer_round_base + rounds published by rank j
    unsigned int peer_round_base;  // grows with every ptfnn_init_chains: the flag domain is never reset (a re-initialised
                                   // handle must not see the previous run's counts as "already published")
    int *smsp_load;                // [num_SMs] ticket counter used to spread serial (SGD) warps over the SM sub-partitions
    // ---- liveness: every wait has a limit on time WITHOUT progress (WaitClock); a failed wait ends the kernel
    long long wait_limit;          // cycles
    unsigned int *heartbeat;       // this rank's progress word (= peer_flags[rank] + kMaxPeers): bumped after every MCMC step
    // ---- co-residency probe (kernels that allocate TMEM are launched without the cooperative-launch proof):
    //      probe = 1: every CTA sets itself up as for a real launch, arrives on probe_count[0], waits (bounded) for
    //      the whole grid and leaves; the host reads how many arrived TOGETHER (probe_count[1] = max seen)
    int probe;
    unsigned int *probe_count;
    // ---- opt-in swap rule of the reference's drafts; 0 = the reference's rule (R:674)
    int swap_kind;
    const double *temperature_global;   // [Rg] (swap_kind != 0 only)
};
constexpr int kSpecWin = 32;       // most steps one speculative window covers
constexpr int kMaxSpecK = 16;      // most CTAs per temperature
constexpr int kSpecWords = 64;     // control words per temperature: kSpecWin step flags, then kMaxSpecK CTA flags
constexpr int kSpecRwRun = 8;      // random-walk steps one CTA takes in a row (a Langevin step costs 7-13 of them)
constexpr int kRowTail = 2;        // words behind the P weights of a published row: eta as raw fp64 bits (R:430 moves the float64)

__device__ __forceinline__ bool swap_due(int rule, int s, int i) {
    return rule == 0 ? (i % s == 0 && i != 0) : ((i + 1) % s == 0);                // R:427 | C:438
}
__device__ __forceinline__ int next_swap_step(int rule, int s, int i) {
    if (rule == 0) { const int k = (i + s - 1) / s; return (k < 1 ? 1 : k) * s; }
    return ((i + s) / s) * s - 1;
}

template <int I, int H, int O>
struct NetSizes {
    static constexpr int P = I * H + H * O + H + O;
    static constexpr int IP = IPad<I>::value;
};

// dynamic shared memory layout (floats unless noted); host computes the same with chain_smem_bytes()
struct ChainSmem {
    int P4;          // P rounded up to 4
    size_t off_w, off_prop, off_gd, off_pgd, off_lw, off_red, off_bar, off_tiles, off_sweep, off_team, off_stage, total;
};
__host__ __device__ inline ChainSmem chain_smem_layout(int P, int IP, int nt, int Rg, bool staged, int n_train,
                                                       int n_test, int team_floats = 0, int lik_floats = 0,
                                                       bool gd_in_smem = true, int tc_bytes = 0, int tc_alias_off = 0) {
    ChainSmem L;
    L.P4 = (P + 3) & ~3;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~(size_t)15; return r; };
    // K5 topologies (tc_bytes > 0): the state vector stays in its global row, the likelihood layout is
    // replaced by the tcgen05 operands, and the SGD pass's TMA tiles + team scratch alias the A tile
    // (the two passes never overlap in time)
    const bool tc = tc_bytes > 0;
    L.off_w = take(tc ? 0 : L.P4 * 4); L.off_prop = take(tc ? 0 : L.P4 * 4);
    // langevin_gradient(w) and langevin_gradient(w_prop): shared memory, or (wide nets) the replica's global rows
    L.off_gd = take(gd_in_smem ? L.P4 * 4 : 0); L.off_pgd = take(gd_in_smem ? L.P4 * 4 : 0);
    L.off_lw = take(tc ? (size_t)tc_bytes : (size_t)lik_floats * 4);   // likelihood layout of the proposal (LikLayout) | tc::Smem
    L.off_red = take((size_t)8 * (nt / 32) * 8);
    L.off_bar = take(8 * 4);
    const size_t tile_bytes = staged ? 0 : (size_t)2 * (kTileRows * IP * 4 + kTileRows * 4);
    L.off_tiles = tc ? L.off_lw + tc_alias_off : take(tile_bytes);
    L.off_sweep = take(Rg > 1 ? (size_t)kSweepChunk * (8 + 8 + 4) : 0);   // one chunk of lhood fields, log(2u) and u for the streaming sweep
    L.off_team = tc ? L.off_tiles + ((tile_bytes + 15) & ~(size_t)15) : take((size_t)team_floats * 4);
    auto pad4 = [](int n) { return (size_t)((n + 3) & ~3); };
    L.off_stage = take(staged ? ((size_t)n_train * IP + pad4(n_train) + (size_t)n_test * IP + pad4(n_test)) * 4 : 0);
    L.total = o;
    return L;
}

// K4: the coordinator's swap round (R:741-752) inside the chain kernel, after the grid barrier.
// Every CTA runs the same sequential sweep (thread 0; lhood fields staged chunk-wise), notes the
// origin of the vectors that end in ITS slots, then pulls (w, eta) of those from the published rows
// (R:435-437: only w and eta are taken back, likelihood / prior stay stale -- SURVEY Q7).
template <int NT>
__device__ __forceinline__ void chain_sweep(const ChainParams &p, int round, int parity, bool apply, double *s_chunk,
                                            int nblocks, int bid) {
    // bid / nblocks: this CTA's index among the CTAs that OWN temperatures (with speculative windows
    // only the first CTA of each group does)
    const int tid = threadIdx.x;
    const int P = p.P;
    const double *L = p.pub_lhood + (size_t)parity * p.Rg;
    const bool log_it = bid == 0 && round < p.max_rounds;
    uint8_t *lg_out = log_it ? p.swap_log + (size_t)round * (p.Rg - 1) : nullptr;
    const float *ur = p.replay ? p.u_swap + (size_t)(round - p.round_begin) * (p.Rg - 1) : nullptr;
    double cur_l = 0.0;
    int cur_src = 0, ns = 0;
    // highest slot this CTA owns: no need to scan past it (block 0 scans everything for the log)
    int last_needed = p.Rg - 1;
    if (bid != 0 && apply) {
        int r_hi = bid;
        while (r_hi + nblocks < p.R) r_hi += nblocks;
        last_needed = p.replica_offset + r_hi;
    }
    auto mine = [&](int slot) {
        const int r = slot - p.replica_offset;
        return apply && r >= 0 && r < p.R && (r % nblocks) == bid;
    };
    double *s_lu = s_chunk + kSweepChunk;                          // log(2 u) of the chunk's pairs
    float *s_u = reinterpret_cast<float *>(s_lu + kSweepChunk);    // their uniforms
    for (int base = 0; base <= last_needed; base += kSweepChunk - 1) {
        // chunk holds original lhood of slots [base, base + kSweepChunk) and the draws of pairs [base, ...)
        __syncthreads();
        for (int k = tid; k < kSweepChunk && base + k < p.Rg; k += NT) {
            s_chunk[k] = __ldcg(&L[base + k]);
            float u = 0.0f;
            if (base + k < p.Rg - 1) {
                if (ur) u = ur[base + k];
                else {
                    uint32_t c[4];
                    philox_draw(p.seed, (uint32_t)round, (uint32_t)(base + k), 0u, kTagSwap, c);
                    u = u01_open_right(c[0]);
                }
            }
            s_u[k] = u;
            s_lu[k] = log(2.0 * (double)u);
        }
        __syncthreads();
        if (tid == 0) {
            if (base == 0) { cur_l = s_chunk[0]; cur_src = 0; }
            const int k_end = min(base + kSweepChunk - 1, last_needed + 1);
            swap_sweep_stream(
                base, k_end, p.Rg, cur_l, cur_src, ns, [&](int k) { return s_chunk[k - base]; },
                [&](int k) { return s_u[k - base]; }, [&](int k) { return s_lu[k - base]; },
                [&](int slot, int origin) { if (mine(slot)) p.swap_src[slot - p.replica_offset] = origin; },
                [&](int k, bool sw) { if (lg_out) lg_out[k] = (uint8_t)sw; },
                p.swap_kind, [&](int slot) { return p.temperature_global[slot]; });
        }
    }
    if (tid == 0) {
        if (last_needed == p.Rg - 1 && mine(p.Rg - 1)) p.swap_src[p.Rg - 1 - p.replica_offset] = cur_src;
        if (bid == 0) { p.swap_counters[0] += ns; p.swap_counters[1] += p.Rg - 1; }
    }
    __syncthreads();
    if (!apply) return;
    for (int r = bid; r < p.R; r += nblocks) {
        const int gsrc = p.swap_src[r];                       // global slot the vector comes from
        if (gsrc != p.replica_offset + r) {
            const int q = p.n_ranks > 1 ? gsrc / p.R : 0;     // equal contiguous blocks: owner rank of that slot
            const int lsrc = gsrc - q * p.R - (p.n_ranks > 1 ? 0 : p.replica_offset);
            if (p.n_ranks > 1 && q != p.rank) {               // the row lives on another GPU: pull it over NVLink
                const float *row = p.peer_rows[q] + ((size_t)parity * p.R + lsrc) * (P + kRowTail);
                for (int j = tid; j < P; j += NT) p.w[(size_t)r * P + j] = ld_sys_f32(&row[j]);
                if (tid == 0) {
                    const unsigned int *t = reinterpret_cast<const unsigned int *>(row + P);
                    p.eta[r] = f64_from_words(ld_sys_u32(&t[0]), ld_sys_u32(&t[1]));
                    p.gd_valid[r] = 0;
                }
            } else {
                const float *row = p.pub_rows + ((size_t)parity * p.R + lsrc) * (P + kRowTail);
                for (int j = tid; j < P; j += NT) p.w[(size_t)r * P + j] = __ldcg(&row[j]);
                if (tid == 0) {
                    const unsigned int *t = reinterpret_cast<const unsigned int *>(row + P);
                    p.eta[r] = f64_from_words(__ldcg(&t[0]), __ldcg(&t[1]));
                    p.gd_valid[r] = 0;
                }
            }
        }
    }
    __syncthreads();
}

// Multi-GPU hand-shake of one swap round (replaces the reference's queues + events across processes,
// R:427-437 / R:730-752, and a host round trip): after the local grid barrier (every local temperature
// has published), block 0 PUSHES this rank's lhood fields into every peer's pub_lhood (8 bytes per
// temperature over NVLink), raises its flag on every peer, and every CTA waits until the flags of all
// ranks have reached this round.  The sweep then reads only local memory; (w, eta) rows are PULLED from
// the owner's pub_rows by the CTA that needs them (in expectation the rows next to the rank boundaries).
// Returns false (in every thread) when a peer never published: the caller leaves the kernel.
template <int NT>
__device__ __forceinline__ bool peer_exchange_lhood(const ChainParams &p, int round, int parity, GridBarrier *bar) {
    const int tid = threadIdx.x;
    const unsigned int want = p.peer_round_base + (unsigned int)(round + 1);
    if (blockIdx.x == 0) {
        const double *mine = p.pub_lhood + (size_t)parity * p.Rg + p.replica_offset;
        for (int q = 0; q < p.n_ranks; ++q) {
            if (q == p.rank) continue;
            double *dst = p.peer_lhood[q] + (size_t)parity * p.Rg + p.replica_offset;
            for (int k = tid; k < p.R; k += NT) dst[k] = __ldcg(&mine[k]);
        }
        __threadfence_system();
        __syncthreads();
        if (tid < p.n_ranks) st_release_sys_u32(&p.peer_flags[tid][p.rank], want);
    }
    int bad = 0;
    if (tid == 0) {
        const unsigned int *flags = p.peer_flags[p.rank];
        for (int q = 0; q < p.n_ranks && !bad; ++q) {
            // progress of rank q = its heartbeat word, read over NVLink only when the limit has run out
            long long t0 = clock64();
            unsigned int last = 0u;
            bool have_last = false;
            while ((int)(ld_acquire_sys_u32(&flags[q]) - want) < 0) {
                __nanosleep(64);
                if (ld_relaxed_u32(&bar->failed)) { bad = 1; break; }
                const long long now = clock64();
                if (now - t0 > p.wait_limit) {
                    const unsigned int beat = ld_sys_u32(&p.peer_flags[q][kMaxPeers]);
                    if (!have_last || beat != last) { last = beat; have_last = true; t0 = now; continue; }
                    atomicExch(&bar->failed, 1u); bad = 1; break;
                }
            }
        }
        __threadfence_system();
    }
    return __syncthreads_or(bad) == 0;
}

// SPEC_T: the instantiation with speculative windows (small ladders); the plain one carries none of their
// state (at 72 registers per thread for 1024 co-resident temperatures every live value counts).
template <int I, int H, int O, int TASK, int NT, bool SPEC_T>
__device__ __forceinline__ void chain_body(const ChainParams &p) {
    constexpr int P = NetSizes<I, H, O>::P;
    constexpr int IP = NetSizes<I, H, O>::IP;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool TEAM = UseSgdTeam<H>::value;
    constexpr bool TC = UseTc<I, H, O, NT>::value;              // K5: tcgen05 likelihood pass
    static_assert(!TC || TEAM, "the tcgen05 path belongs to the wide-hidden team topologies");
    const ChainSmem L = chain_smem_layout(P, IP, NT, p.external_swap ? 1 : p.Rg, p.staged != 0, p.train.n, p.test.n, TEAM ? team_smem_floats(H, O) : 0,
                                          H * LikLayout<I, O>::LW, !TEAM, TC ? tc::Smem<I, H, O>::total : 0, tc::Smem<I, H, O>::off_a);
    unsigned char *s_tc = smem_raw + L.off_lw;
    float *s_lw = reinterpret_cast<float *>(smem_raw + L.off_lw);
    float *s_team = reinterpret_cast<float *>(smem_raw + L.off_team);
    float *s_w = reinterpret_cast<float *>(smem_raw + L.off_w);
    float *s_prop = reinterpret_cast<float *>(smem_raw + L.off_prop);
    float *s_gd = reinterpret_cast<float *>(smem_raw + L.off_gd);      // re-pointed per replica when TEAM
    float *s_pgd = reinterpret_cast<float *>(smem_raw + L.off_pgd);
    double *s_red = reinterpret_cast<double *>(smem_raw + L.off_red);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + L.off_bar);
    double *s_sweep = reinterpret_cast<double *>(smem_raw + L.off_sweep);
    __shared__ int s_flag[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time setup: mbarriers; stage both datasets into shared memory with TMA bulk copies
    DataView train = p.train, test = p.test;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_init(&s_bar[2], 1);
        mbar_fence_init();
    }
    __syncthreads();
    SgdStream stream;
    if (p.staged) {
        auto pad4 = [](int n) { return (n + 3) & ~3; };
        float *sx_tr = reinterpret_cast<float *>(smem_raw + L.off_stage);
        float *sy_tr = sx_tr + (size_t)train.n * IP;
        float *sx_te = sy_tr + pad4(train.n);
        float *sy_te = sx_te + (size_t)test.n * IP;
        if (tid == 0) {
            const uint32_t b0 = (uint32_t)train.n * IP * 4u, b1 = (uint32_t)pad4(train.n) * 4u;
            const uint32_t b2 = (uint32_t)test.n * IP * 4u, b3 = (uint32_t)pad4(test.n) * 4u;
            mbar_arrive_expect_tx(&s_bar[2], b0 + b1 + b2 + b3);
            tma_load_1d(sx_tr, train.x, b0, &s_bar[2]);
            tma_load_1d(sy_tr, train.y, b1, &s_bar[2]);
            tma_load_1d(sx_te, test.x, b2, &s_bar[2]);
            tma_load_1d(sy_te, test.y, b3, &s_bar[2]);
        }
        mbar_wait(&s_bar[2], 0);
        train.x = sx_tr; train.y = sy_tr; test.x = sx_te; test.y = sy_te;
    } else {
        float *tiles = reinterpret_cast<float *>(smem_raw + L.off_tiles);
        stream.tile_x0 = tiles;
        stream.tile_x1 = tiles + kTileRows * IP;
        stream.tile_y0 = tiles + 2 * kTileRows * IP;
        stream.tile_y1 = tiles + 2 * kTileRows * IP + kTileRows;
        stream.bar0 = &s_bar[0]; stream.bar1 = &s_bar[1];
    }
    stream.parity0 = stream.parity1 = 0u;
    tc::State tcst;
    if constexpr (TC) tc::setup<I, H, O>(s_tc, tcst);            // TMEM: 256 columns per CTA, two CTAs per SM
    // ---- liveness.  A launch on a handle whose earlier wait failed does nothing; the co-residency probe
    //      (TMEM kernels are launched without the cooperative-launch proof) counts the CTAs that are resident
    //      TOGETHER with everything a real launch holds (registers, shared memory, TMEM columns) and leaves.
    {
        int dead = 0;
        if (tid == 0) {
            dead = (int)ld_relaxed_u32(&p.barrier->failed);
            if (p.probe && !dead) {
                const unsigned int n = atomicAdd(&p.probe_count[0], 1u) + 1u;
                atomicMax(&p.probe_count[1], n);
                const long long t0 = clock64();
                while (ld_relaxed_u32(&p.probe_count[0]) < gridDim.x && clock64() - t0 < (1ll << 22)) __nanosleep(64);   // ~2 ms
                atomicMax(&p.probe_count[1], ld_relaxed_u32(&p.probe_count[0]));
                __nanosleep(2000);                     // late arrivals of a resident grid are counted by everybody
                atomicSub(&p.probe_count[0], 1u);      // a CTA that starts after this one has left is not co-resident with it
                dead = 1;
            }
        }
        if (__syncthreads_or(dead)) {
            if constexpr (TC) tc::teardown<I, H, O>(tcst);
            return;
        }
    }
    // likelihood sums of weight vector wv on both data sets (K5 path)
    auto tc_likelihood = [&](const float *wv, bool with_test, double &a0, double &a1, int &c0, double &b0, double &b1, int &c1) {
        if constexpr (TC) {
            tc::build_b<I, H, O>(s_tc, wv, tid, NT);
            tc::fence_async_smem();
            __syncthreads();
            tc::lik_pass<I, H, O, TASK, NT, false>(s_tc, tcst, p.a_train, train.y, train.n, wv, a0, a1, c0, nullptr, nullptr);
            if (with_test)
                tc::lik_pass<I, H, O, TASK, NT, false>(s_tc, tcst, p.a_test, test.y, test.n, wv, b0, b1, c1, nullptr, nullptr);
        }
    };

    // ---- which warp runs the serial recurrence.  Warps are bound to one of the SM's four
    // sub-partitions (hardware warp slot % 4); co-resident CTAs would otherwise all put their
    // serial warp on the same sub-partition and queue for its issue port (measured: 2x slower rows).
    __shared__ unsigned int s_hw[NW];
    if (lane == 0) s_hw[warp] = hw_warpid() & 3u;
    __syncthreads();
    const unsigned int smid = hw_smid();
    if (tid == 0) {
        // race-free: a per-SM ticket; ticket % 4 is the sub-partition this CTA's serial warp should use
        const unsigned int target = (unsigned int)atomicAdd(&p.smsp_load[smid], 1) & 3u;
        int best = 0;
        for (int w = NW - 1; w >= 0; --w)
            if (s_hw[w] == target) best = w;
        s_flag[0] = best;
    }
    __syncthreads();
    const int sgd_warp = s_flag[0];
    const bool is_sgd_warp = warp == sgd_warp;
    const int lik_tid = (warp < sgd_warp ? tid : tid - 32);     // index inside the team of the other warps

    __shared__ int s_plan[SPEC_T ? kSpecWin + 1 : 1];      // owner CTA of each step of the window; [kSpecWin] = its length
    __shared__ int s_res[2];                                // the window's first accepted step; first CTA holding the base gradient
    const int nblocks = gridDim.x;
    // speculative windows: K CTAs per temperature (host guarantees gridDim.x == R * K and co-residency)
    const int K = (SPEC_T && !TEAM && p.spec_k > 1) ? p.spec_k : 1;
    const bool SPEC = SPEC_T && K > 1;
    const int kq = SPEC ? (int)blockIdx.x % K : 0;               // position inside the window
    const int vblock = SPEC ? (int)blockIdx.x / K : (int)blockIdx.x, nvb = SPEC ? nblocks / K : nblocks;
    int step = p.step_begin;
    int round = p.round_begin;
    const double inv_sig2 = 1.0 / ((double)p.step_w * (double)p.step_w);

    while (step < p.step_end) {
        int seg_last = p.Rg > 1 ? next_swap_step(p.swap_rule, p.swap_interval, step) : p.step_end - 1;
        bool swap_at_end = p.Rg > 1;
        if (seg_last > p.step_end - 1) { seg_last = p.step_end - 1; swap_at_end = false; }
        const int parity = round & 1;

        for (int r = vblock; r < p.R; r += nvb) {
            // ---------------- load this replica's state ----------------
            if constexpr (TEAM) { s_gd = p.gd_cache + (size_t)r * P; s_pgd = p.pgd_buf + (size_t)r * P; }
            if constexpr (TC) { s_w = p.w + (size_t)r * P; s_prop = p.prop_buf + (size_t)r * P; }   // state and proposal stay in global rows
            double eta, tau, lik, prior_cur, last_rtr, last_rte, last_atr, last_ate;
            int n_acc, init_count, gd_valid;
            auto load_state = [&]() {            // __ldcg: with speculative windows another CTA wrote it
                for (int j = tid; j < P; j += NT) {
                    if constexpr (!TC) s_w[j] = __ldcg(&p.w[(size_t)r * P + j]);
                    if constexpr (!TEAM) s_gd[j] = __ldcg(&p.gd_cache[(size_t)r * P + j]);
                }
                eta = __ldcg(&p.eta[r]); tau = __ldcg(&p.tau[r]); lik = __ldcg(&p.lik[r]); prior_cur = __ldcg(&p.prior[r]);
                n_acc = __ldcg(&p.n_acc[r]); init_count = __ldcg(&p.init_count[r]);
                gd_valid = p.memo ? __ldcg(&p.gd_valid[r]) : 0;
                last_rtr = __ldcg(&p.last4[r * 4 + 0]); last_rte = __ldcg(&p.last4[r * 4 + 1]);
                last_atr = __ldcg(&p.last4[r * 4 + 2]); last_ate = __ldcg(&p.last4[r * 4 + 3]);
                __syncthreads();
            };
            auto store_state = [&]() {
                for (int j = tid; j < P; j += NT) {
                    if constexpr (!TC) p.w[(size_t)r * P + j] = s_w[j];
                    if constexpr (!TEAM) { if (p.memo && gd_valid > 0) p.gd_cache[(size_t)r * P + j] = s_gd[j]; }
                }
                if (tid == 0) {
                    p.eta[r] = eta; p.tau[r] = tau; p.lik[r] = lik; p.prior[r] = prior_cur;
                    p.n_acc[r] = n_acc; p.init_count[r] = init_count; p.gd_valid[r] = gd_valid != 0;
                    p.last4[r * 4 + 0] = last_rtr; p.last4[r * 4 + 1] = last_rte;
                    p.last4[r * 4 + 2] = last_atr; p.last4[r * 4 + 3] = last_ate;
                }
                __syncthreads();
            };
            load_state();
            const double temperature = p.temperature[r];
            const uint32_t gr = (uint32_t)(p.replica_offset + r);
            const uint32_t rng_stream = p.crn ? kStreamCommon : gr;

            // One iteration = one step, or (speculative windows) W consecutive steps shared out among the K
            // CTAs of the group, each step evaluated from the state at the window base, i.e. as if the earlier
            // steps of the window were rejected.  The first accepted step k* makes steps 0..k* stand; its CTA
            // installs the new state and the window after it starts at base + k* + 1.  Results are those of
            // the sequential chain bit for bit: the draws are indexed by the step, a rejected step leaves
            // nothing behind but its trace row and tau, and trace rows past k* are rewritten later.
            // The window is PACKED by cost: every CTA takes one Langevin step (two SGD epochs) and the
            // random-walk steps right before it -- several times cheaper -- so one window of little more than
            // one Langevin step's duration covers ~K / l_prob steps instead of K.  Every CTA of the group
            // derives the same plan from the lx draws of the steps ahead.
            int ibase = step;
            while (ibase <= seg_last) {
                int W = 1;
                if (SPEC) {
                    if (ibase != step) load_state();
                    int wcap = min(kSpecWin, seg_last - ibase + 1);
                    // the temperature switch (a11) changes the state whatever the MH outcome: the step that
                    // performs it opens a window of its own and no window runs across it
                    const int sw = (int)p.pt_samples;
                    if (init_count == 0 && (double)sw == p.pt_samples) {
                        if (ibase == sw) wcap = 1;
                        else if (ibase < sw && ibase + wcap > sw) wcap = sw - ibase;
                    }
                    // warp 0 plans.  Two ways to share the steps out, both with one Langevin step per CTA at most:
                    //   "apart":  the j-th Langevin step goes to CTA j, the random-walk steps to CTAs K-1, K-2, ...
                    //             (kSpecRwRun to a CTA): nothing runs before a Langevin step, the window lasts one of them;
                    //             it ends before the step at which the two ranges would meet.
                    //   "riding": a step goes to the CTA whose index is the number of Langevin steps before it, i.e. a CTA
                    //             evaluates the random-walk steps that precede "its" Langevin step and then that step (a run
                    //             of more than kSpecRwRun random-walk steps moves on): K Langevin steps per window, each
                    //             a few random-walk steps late.
                    // "apart" is used unless "riding" covers more steps (with few CTAs per temperature it does).  Measured on one
                    // box, apart | riding: Sunspot (K = 14) 737 k | 660 k replica-steps/s; 10 steps of 128 temperatures (K = 8)
                    // 11.7 | 12.6 ms, of 256 (K = 4) 20.0 | 19.3 ms, of 512 (K = 2) 39.3 | 33.9 ms.
                    if (tid < 32) {
                        bool lgt = false;
                        if (tid < wcap) {
                            const int it = ibase + tid;
                            float lxt;
                            if (p.replay) lxt = p.lx[(size_t)r * p.replay_n + (it - p.step_begin)];
                            else lxt = philox_step_scalars(p.seed, (uint32_t)it, p.crn ? kStreamCommon : (uint32_t)(p.replica_offset + r), (uint32_t)(p.replica_offset + r)).lx;
                            lgt = p.use_lg && ((double)lxt < p.l_prob);
                        }
                        const unsigned int in_cap = wcap >= 32 ? 0xffffffffu : ((1u << wcap) - 1u);
                        const unsigned int lgm = __ballot_sync(0xffffffffu, lgt) & in_cap;
                        const unsigned int before = (1u << tid) - 1u;
                        const int n_before = __popc(lgm & before), m_before = __popc(~lgm & in_cap & before);
                        const int n_incl = n_before + (lgt ? 1 : 0), m_incl = m_before + (lgt ? 0 : 1);
                        const int owner_apart = lgt ? n_before : K - 1 - m_before / kSpecRwRun;
                        const int owner_riding = p.use_lg ? max(n_before, tid / (kSpecRwRun + 1)) : tid;      // (non-decreasing in the step)
                        const int w_apart = __popc(__ballot_sync(0xffffffffu, tid < wcap && n_incl + (m_incl + kSpecRwRun - 1) / kSpecRwRun <= K));
                        const int w_riding = __popc(__ballot_sync(0xffffffffu, tid < wcap && owner_riding < K));
                        const bool apart = p.use_lg && (p.spec_plan == 1 || (p.spec_plan == 0 && w_apart + (K >= 8 ? 2 : 0) >= w_riding));
                        s_plan[tid] = apart ? owner_apart : owner_riding;
                        if (tid == 0) s_plan[kSpecWin] = apart ? w_apart : w_riding;
                    }
                    __syncthreads();
                    W = s_plan[kSpecWin];
                }
                const int gd_valid0 = gd_valid;      // langevin_gradient(w) known at the window base?
                bool accept = false;
                for (int t = 0; t < W && !accept; ++t) {
                if (SPEC && s_plan[t] != kq) continue;
                const int i = ibase + t;
                // ---- a11: temperature schedule inside the chain (R:317-324)
                double adapt = init_count ? 1.0 : temperature;
                if ((double)i == p.pt_samples && init_count == 0) {
                    adapt = 1.0;
                    init_count = 1;
                    double s[2] = {0.0, 0.0};
                    double dummy = 0.0;
                    int c0 = 0;
                    if constexpr (TC) {
                        double d2 = 0.0, d3 = 0.0;
                        int c1 = 0;
                        tc_likelihood(s_w, false, s[0], dummy, c0, d2, d3, c1);
                    } else {
                        lik_prepare<I, H, O>(s_lw, s_w, tid, NT);
                        __syncthreads();
                        lik_fast<I, H, O, TASK>(s_lw, s_w, train, tid, NT, s[0], dummy, c0);
                    }
                    block_sum<2, NT>(s, s_red);
                    if constexpr (TASK == kTaskReg)
                        lik = (-0.5 * train.n * log(2.0 * 3.14159265358979323846 * tau) - 0.5 * s[0] / tau) / adapt;
                    else
                        lik = s[0] / adapt;
                }
                // ---- draws (a9): lx, proposal noise, eta noise, MH uniform
                float lx, z_eta, u;
                const int di = i - p.step_begin;
                if (p.replay) {
                    const size_t q = (size_t)r * p.replay_n + di;
                    lx = p.lx[q];
                    z_eta = (TASK == kTaskReg && p.z_eta) ? p.z_eta[q] : 0.0f;
                    u = p.u[q];
                } else {
                    const StepDraws sd = philox_step_scalars(p.seed, (uint32_t)i, rng_stream, gr);
                    lx = sd.lx; z_eta = sd.z_eta; u = sd.u;
                }
                const bool lg = p.use_lg && ((double)lx < p.l_prob);              // R:329
                // ---- Langevin branch, first SGD epoch: w_gd = langevin_gradient(w)   (R:330)
                if (lg && !gd_valid) {
                    if constexpr (TEAM) { if (tid < kTeamThreads) sgd_pass_team<I, H, O, TASK, kTeamThreads>(s_w, s_gd, train, p.staged != 0, p.lr, stream, s_team); }
                    else if (is_sgd_warp) sgd_pass<I, H, O, TASK>(s_w, s_gd, train, p.staged != 0, p.lr, stream);
                    __syncthreads();
                    gd_valid = p.memo;
                }
                // ---- proposal: w_prop = (w_gd | w) + step_w * z                     (R:331 | R:353)
                {
                    const float *base = lg ? s_gd : s_w;
                    if (p.replay) {
                        const float *zz = p.z + ((size_t)r * p.replay_n + di) * P;
                        for (int j = tid; j < P; j += NT) {
                            const float v = fmaf(p.step_w, zz[j], base[j]);
                            s_prop[j] = v;
                            if constexpr (!TC) lik_scatter<I, H, O>(s_lw, j, v);
                        }
                    } else {
                        for (int b = tid; b < (P + 3) / 4; b += NT) {
                            float z4[4];
                            philox_step_normals4(p.seed, (uint32_t)i, rng_stream, (uint32_t)b, z4);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int j = 4 * b + k;
                                if (j < P) {
                                    const float v = fmaf(p.step_w, z4[k], base[j]);
                                    s_prop[j] = v;
                                    if constexpr (!TC) lik_scatter<I, H, O>(s_lw, j, v);
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                // ---- second SGD epoch on warp 0 (R:332) WHILE the other warps evaluate the
                //      proposal's likelihood on train and test (R:360-362)
                double s[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] = 0.0;
                int c_tr = 0, c_te = 0;
                // The test-set pass only feeds rmse_test / acc_test, which are recorded on ACCEPTANCE (R:362,
                // R:399-406): with `memo` (skip work whose result is provably unused, like the gradient memo) it
                // runs after the MH decision, for accepted proposals only.  The overlapped pass below keeps it:
                // there it hides behind the SGD epoch.
                const bool lazy_test = p.memo != 0 && !(lg && NW > 1 && !TEAM);
                if (lg && NW > 1 && !TEAM) {
                    // (measured alternatives: fewer likelihood warps make that pass the critical path; running
                    // the two one after the other costs the same as overlapping them -- DESIGN section 5)
                    if (is_sgd_warp) {
                        sgd_pass<I, H, O, TASK>(s_prop, s_pgd, train, p.staged != 0, p.lr, stream);
                    } else {
                        lik_fast<I, H, O, TASK>(s_lw, s_prop, train, lik_tid, NT - 32, s[0], s[1], c_tr);
                        lik_fast<I, H, O, TASK>(s_lw, s_prop, test, lik_tid, NT - 32, s[2], s[3], c_te);
                    }
                } else {
                    if (lg) {
                        if constexpr (TEAM) { if (tid < kTeamThreads) sgd_pass_team<I, H, O, TASK, kTeamThreads>(s_prop, s_pgd, train, p.staged != 0, p.lr, stream, s_team); }
                        else if (is_sgd_warp) sgd_pass<I, H, O, TASK>(s_prop, s_pgd, train, p.staged != 0, p.lr, stream);
                        __syncthreads();
                    }
                    if constexpr (TC) {
                        tc_likelihood(s_prop, !lazy_test, s[0], s[1], c_tr, s[2], s[3], c_te);
                    } else {
                        lik_fast<I, H, O, TASK>(s_lw, s_prop, train, tid, NT, s[0], s[1], c_tr);
                        if (!lazy_test) lik_fast<I, H, O, TASK>(s_lw, s_prop, test, tid, NT, s[2], s[3], c_te);
                    }
                }
                __syncthreads();
                // ---- reductions: likelihood sums, |w_prop|^2 (prior), Langevin asymmetry norms
                s[4] = (double)c_tr; s[5] = (double)c_te;
                {
                    double sq = 0.0, first = 0.0, second = 0.0;
                    for (int j = tid; j < P; j += NT) {
                        const float wp = s_prop[j];
                        sq += (double)wp * (double)wp;
                        if (lg) {
                            const double a = (double)s_w[j] - (double)s_pgd[j];     // wc_delta (R:336)
                            const double b = (double)wp - (double)s_gd[j];          // wp_delta (R:337)
                            first += a * a; second += b * b;
                        }
                    }
                    s[6] = sq;
                    s[7] = second - first;   // -> diff_prop = 0.5*(second-first)/sigma^2 / adapttemp
                }
                block_sum<8, NT>(s, s_red);
                // ---- scalars (every thread computes the same values; thread 0 records them)
                double eta_pro = eta;
                if constexpr (TASK == kTaskReg) {
                    eta_pro = eta + (double)p.step_eta * (double)z_eta;             // R:355
                    tau = exp(eta_pro);                                             // R:356
                }
                double lik_prop, rmse_tr, rmse_te, acc_tr = 0.0, acc_te = 0.0, prior_prop;
                if constexpr (TASK == kTaskReg) {
                    lik_prop = (-0.5 * train.n * log(2.0 * 3.14159265358979323846 * tau) - 0.5 * s[0] / tau) / adapt;  // R:204-205
                    rmse_tr = sqrt(s[0] / train.n);
                    rmse_te = sqrt(s[2] / test.n);
                    prior_prop = -((I * H + H + 2) / 2.0) * log(p.sigma_sq) - s[6] / (2.0 * p.sigma_sq) -
                                 (1.0 + p.nu1) * log(tau) - p.nu2 / tau;           // R:218-220
                } else {
                    lik_prop = s[0] / adapt;                                        // C:222
                    rmse_tr = sqrt(s[1] / train.n);
                    rmse_te = sqrt(s[3] / test.n);
                    acc_tr = 100.0 * (s[4] / train.n);
                    acc_te = 100.0 * (s[5] / test.n);
                    prior_prop = -((I * H + H + O + H * O) / 2.0) * log(p.sigma_sq) - s[6] / (2.0 * p.sigma_sq);  // C:227-229
                }
                const double diff_prop = lg ? (0.5 * s[7] * inv_sig2) / adapt : 0.0;   // R:341-346 (Q4)
                const double a = (lik_prop - lik) + (prior_prop - prior_cur) + diff_prop;   // R:365-373
                double mh = exp(a);
                if (!(mh < 1.0)) mh = 1.0;     // min(1, .): overflow -> 1 (R:375); NaN -> 1 (Python min)
                accept = (double)u < mh;                                              // R:395
                if (lazy_test && accept) {                                            // R:362 for the proposals that need it
                    double t[3] = {0.0, 0.0, 0.0};
                    int ct = 0;
                    if constexpr (TC) tc::lik_pass<I, H, O, TASK, NT, false>(s_tc, tcst, p.a_test, test.y, test.n, s_prop, t[0], t[1], ct, nullptr, nullptr);
                    else lik_fast<I, H, O, TASK>(s_lw, s_prop, test, tid, NT, t[0], t[1], ct);
                    t[2] = (double)ct;
                    block_sum<3, NT>(t, s_red);
                    if constexpr (TASK == kTaskReg) rmse_te = sqrt(t[0] / test.n);
                    else { rmse_te = sqrt(t[1] / test.n); acc_te = 100.0 * (t[2] / test.n); }
                }
                // ---- traces of row i+1
                const size_t ti = (size_t)r * p.S + (i + 1);
                if (tid == 0) {
                    p.accept_list[ti] = n_acc;                                        // R:380 (count BEFORE)
                    p.lik_prop[ti] = (TASK == kTaskReg) ? lik_prop : lik_prop * adapt;    // R:391 / C:404
                    if (p.debug) {
                        p.dbg_prior[ti] = prior_prop; p.dbg_diff[ti] = diff_prop; p.dbg_mh[ti] = mh;
                        p.dbg_acc[ti] = accept ? 1 : 0;
                    }
                }
                if (accept) {
                    n_acc += 1; lik = lik_prop; prior_cur = prior_prop; eta = eta_pro;
                    last_rtr = rmse_tr; last_rte = rmse_te;
                    if constexpr (TASK == kTaskCls) { last_atr = acc_tr; last_ate = acc_te; }   // C:414-415 (Q13)
                    else { last_atr = 0.0; last_ate = 0.0; }                                   // R:403-404
                    for (int j = tid; j < P; j += NT) {
                        const float v = s_prop[j];
                        s_w[j] = v;
                        if (lg && p.memo) s_gd[j] = s_pgd[j];   // langevin_gradient(new w) is already known
                    }
                    gd_valid = (lg && p.memo) ? 1 : 0;
                }
                if (tid == 0) {
                    p.rmse_tr[ti] = last_rtr; p.rmse_te[ti] = last_rte;
                    p.acc_tr[ti] = last_atr; p.acc_te[ti] = last_ate;
                }
                __syncthreads();
                float *pw = p.pos_w + ti * P;
                // accepted: the new state (R:408); rejected: the previous row is carried (R:417) -- it was
                // written by these same threads (or is row 0 = ones), so no copy of it is kept on chip
                if (accept) { for (int j = tid; j < P; j += NT) pw[j] = s_w[j]; }
                else {
                    const float *prev = p.pos_w + ((size_t)r * p.S + ibase) * P;      // the row before the window
                    for (int j = tid; j < P; j += NT) pw[j] = __ldcg(&prev[j]);         // (maybe written by another CTA)
                }
                if (SPEC && tid == 0) p.spec_flag[r * kSpecWords + t] = ((unsigned int)(ibase + 1) << 1) | (accept ? 1u : 0u);
                }   // the steps of this CTA
                if (!SPEC) {
                    if (tid == 0) atomicAdd(p.heartbeat, 1u);     // progress, as seen by the waits of other CTAs / ranks
                    ++ibase; continue;
                }
                // ---- resolve the window
                // step flag = (window base + 1) << 1 | accepted;  CTA flag = (window base + 1) << 1 | this CTA
                // computed langevin_gradient(w) of the base state
                const bool made_gd = !accept && p.memo && gd_valid && !gd_valid0;
                if (tid == 0) {
                    p.spec_flag[r * kSpecWords + kSpecWin + kq] = ((unsigned int)(ibase + 1) << 1) | (made_gd ? 1u : 0u);
                    atomicAdd(p.heartbeat, 1u);
                }
                if (!grid_barrier(&p.spec_bar[r], (unsigned int)K, p.wait_limit, p.heartbeat, /*spin=*/true)) goto chain_exit;
                int kstar = W, kgd = K;
                if (tid < 32) {                              // one load per step of the window, first acceptance by ballot
                    const unsigned int want = ((unsigned int)(ibase + 1) << 1) | 1u;
                    const bool hit = tid < W && __ldcg(&p.spec_flag[r * kSpecWords + tid]) == want;
                    const bool hgd = tid < K && __ldcg(&p.spec_flag[r * kSpecWords + kSpecWin + tid]) == want;
                    const unsigned int hm = __ballot_sync(0xffffffffu, hit), gm = __ballot_sync(0xffffffffu, hgd);
                    if (tid == 0) { s_res[0] = hm ? __ffs(hm) - 1 : W; s_res[1] = gm ? __ffs(gm) - 1 : K; }
                }
                __syncthreads();
                kstar = s_res[0]; kgd = s_res[1];
                const int committed = min(kstar, W - 1);     // the last step of the window that stands
                const int owner = s_plan[committed];         // its CTA
                // that CTA holds exactly the chain's state after that step: the accepted vector, or (no
                // acceptance) the unchanged state with the last proposed tau.  Without an acceptance the memo
                // langevin_gradient(w) stays valid for the next window: the first CTA that computed it stores it.
                if (kstar == W && kgd < K) {
                    if (kq == kgd && kq != owner)
                        for (int j = tid; j < P; j += NT) p.gd_cache[(size_t)r * P + j] = s_gd[j];
                    if (kq == owner && !gd_valid) gd_valid = -1;        // valid, but this CTA does not hold the vector
                }
                if (kq == owner) store_state();
                if (!grid_barrier(&p.spec_bar[r], (unsigned int)K, p.wait_limit, p.heartbeat, /*spin=*/true)) goto chain_exit;
                ibase += committed + 1;
            }

            // ---------------- publish for the hand-shake (R:427-431 / C:438-440) ----------------
            if (SPEC) { if (kq != 0) continue; load_state(); }
            if (swap_at_end) {
                float *row = p.pub_rows + ((size_t)parity * p.R + r) * (P + kRowTail);
                for (int j = tid; j < P; j += NT) row[j] = s_w[j];
                if (tid == 0) {
                    put_f64_words(&row[P], eta);
                    p.pub_lhood[(size_t)parity * p.Rg + p.replica_offset + r] =
                        (TASK == kTaskReg) ? lik * temperature : lik;                 // R:430 (Q8) | C:439
                }
            }
            // ---------------- store state ----------------
            if (!SPEC) store_state();
        }
        step = seg_last + 1;

        if (swap_at_end) {
            if (p.external_swap) break;   // multi-GPU: the host completes the round (ptfnn_swap_*)
            // ---------------- K4: swap round ----------------
            // (a wait that gave up means unpublished rows: leave without touching state, traces or the window)
            if (!grid_barrier(p.barrier, nblocks, p.wait_limit, p.heartbeat)) goto chain_exit;
            if (p.n_ranks > 1 && !peer_exchange_lhood<NT>(p, round, parity, p.barrier)) goto chain_exit;
            if (kq == 0) chain_sweep<NT>(p, round, parity, /*apply=*/true, s_sweep, nvb, vblock);
            __syncthreads();
            if (SPEC && !grid_barrier(p.barrier, nblocks, p.wait_limit, p.heartbeat)) goto chain_exit;   // the pulled vectors are visible to every CTA of the group
            ++round;
        }
    }

    // ---------------- left-over coordinator round on the exit vectors (R:442-444) -------------
    if (p.final_round && !p.external_swap && step >= p.step_end) {
        const int parity = round & 1;
        if (kq == 0)
            for (int r = vblock; r < p.R; r += nvb)
                if (tid == 0) p.pub_lhood[(size_t)parity * p.Rg + p.replica_offset + r] = __ldcg(&p.lik[r]);
        if (!grid_barrier(p.barrier, nblocks, p.wait_limit, p.heartbeat)) goto chain_exit;
        if (p.n_ranks > 1 && !peer_exchange_lhood<NT>(p, round, parity, p.barrier)) goto chain_exit;
        if (blockIdx.x == 0) chain_sweep<NT>(p, round, parity, /*apply=*/false, s_sweep, nvb, 0);
    }
chain_exit:
    if constexpr (TC) tc::teardown<I, H, O>(tcst);
}

// The kernel proper.  Registers are bounded through the number of co-resident CTAs per SM (MINB).  (The five-warp
// tcgen05 geometry gets 168 registers at two CTAs per SM: the register file is carved up per PAIR of warps, so 160
// threads cost what 192 do -- a build with 200 registers left one CTA per SM resident, as the probe launch reported.)
template <int I, int H, int O, int TASK, int NT, int MINB, bool SPEC_T = false>
__global__ void __launch_bounds__(NT, MINB) chain_kernel(const ChainParams p) {
    chain_body<I, H, O, TASK, NT, SPEC_T>(p);
}

// ==========================================================================================
// pre-loop part of ptReplica.run (R:266-285 / C:271-284): eta = log var(fx - y), tau, prior,
// current likelihood; also resets the carried trace rows.  One CTA per replica.
// ==========================================================================================
struct InitParams {
    int R, S;
    double sigma_sq, nu1, nu2;
    const double *temperature;
    DataView train, test;
    const float *w;
    double *eta, *tau, *lik, *prior;
    double *init_rmse;   // [R][2] (train, test) -- informational
};

template <int I, int H, int O, int TASK, int NT>
__global__ void __launch_bounds__(NT) init_kernel(const InitParams p) {
    constexpr int P = NetSizes<I, H, O>::P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_w = reinterpret_cast<float *>(smem_raw);
    double *s_red = reinterpret_cast<double *>(smem_raw + (((size_t)P * 4 + 15) & ~(size_t)15));
    const int r = blockIdx.x, tid = threadIdx.x;
    for (int j = tid; j < P; j += NT) s_w[j] = p.w[(size_t)r * P + j];
    __syncthreads();
    double eta = 0.0, tau = 1.0;                                                  // C:263 junk variable
    if constexpr (TASK == kTaskReg) {
        // np.var(pred_train - y_train) (R:270): ONE forward per row; each thread keeps the shifted sums
        // sum(d - d0), sum((d - d0)^2) of its rows in fp64 with d0 = the residual of row 0 (no cancellation:
        // the residuals of an untrained net differ from one another by O(1) like they differ from d0),
        // var = S2/n - (S1/n)^2.  (Round 1 evaluated every row twice for a two-pass variance: 7.5 ms per launch
        // at 1024 temperatures.)
        constexpr int IP = IPad<I>::value;
        auto residual = [&](int row) {
            DataView one{p.train.x + (size_t)row * IP, p.train.y + row, 1};
            float fx = 0.0f;
            double sse = 0.0, d1 = 0.0;
            int c = 0;
            lik_rows_impl<I, H, O, TASK, 1, true, true>(s_w, one, 0, 1, sse, d1, c, &fx, nullptr);
            return (double)fx - (double)p.train.y[row];
        };
        const double d0 = residual(0);
        double sv[2] = {0.0, 0.0};
        for (int row = tid; row < p.train.n; row += NT) {
            const double dv = residual(row) - d0;
            sv[0] += dv; sv[1] += dv * dv;
        }
        block_sum<2, NT>(sv, s_red);
        const double m1 = sv[0] / p.train.n;
        double var = sv[1] / p.train.n - m1 * m1;
        if (var < 0.0) var = 0.0;
        sv[0] = var * p.train.n;
        eta = log(sv[0] / p.train.n);
        tau = exp(eta);                                                           // R:271
    }
    double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    int c0 = 0, c1 = 0;
    lik_rows<I, H, O, TASK, true, false>(s_w, p.train, tid, NT, s[0], s[1], c0);
    lik_rows<I, H, O, TASK, true, false>(s_w, p.test, tid, NT, s[2], s[3], c1);
    for (int j = tid; j < P; j += NT) s[4] += (double)s_w[j] * (double)s_w[j];
    block_sum<5, NT>(s, s_red);
    if (tid == 0) {
        const double T = p.temperature[r];
        double lik, prior;
        if constexpr (TASK == kTaskReg) {
            lik = (-0.5 * p.train.n * log(2.0 * 3.14159265358979323846 * tau) - 0.5 * s[0] / tau) / T;   // R:284
            prior = -((I * H + H + 2) / 2.0) * log(p.sigma_sq) - s[4] / (2.0 * p.sigma_sq) -
                    (1.0 + p.nu1) * log(tau) - p.nu2 / tau;                                              // R:280
            p.init_rmse[r * 2 + 0] = sqrt(s[0] / p.train.n);
            p.init_rmse[r * 2 + 1] = sqrt(s[2] / p.test.n);
        } else {
            lik = s[0] / T;                                                                              // C:283
            prior = -((I * H + H + O + H * O) / 2.0) * log(p.sigma_sq) - s[4] / (2.0 * p.sigma_sq);      // C:281
            p.init_rmse[r * 2 + 0] = sqrt(s[1] / p.train.n);
            p.init_rmse[r * 2 + 1] = sqrt(s[3] / p.test.n);
        }
        p.eta[r] = eta; p.tau[r] = tau; p.lik[r] = lik; p.prior[r] = prior;
    }
}

// ==========================================================================================
// single-operation kernels (one CTA per weight vector: blockIdx.x indexes a batch of them, e.g. the
// posterior samples of ptfnn_op_posterior_predictive)
// ==========================================================================================
template <int I, int H, int O, int TASK, int NT>
__global__ void __launch_bounds__(NT) op_forward_kernel(const float *w, DataView d, float *fx, float *prob,
                                                        double *sums /* s0, s1, correct */) {
    constexpr int P = NetSizes<I, H, O>::P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_w = reinterpret_cast<float *>(smem_raw);
    double *s_red = reinterpret_cast<double *>(smem_raw + (((size_t)P * 4 + 15) & ~(size_t)15));
    w += (size_t)blockIdx.x * P;
    if (fx) fx += (size_t)blockIdx.x * d.n;
    if (prob) prob += (size_t)blockIdx.x * d.n * O;
    sums += (size_t)blockIdx.x * 3;
    for (int j = threadIdx.x; j < P; j += NT) s_w[j] = w[j];
    __syncthreads();
    double s[3] = {0.0, 0.0, 0.0};
    int c = 0;
    lik_rows<I, H, O, TASK, true, true>(s_w, d, threadIdx.x, NT, s[0], s[1], c, fx, prob);
    s[2] = (double)c;
    block_sum<3, NT>(s, s_red);
    if (threadIdx.x == 0) { sums[0] = s[0]; sums[1] = s[1]; sums[2] = s[2]; }
}

template <int I, int H, int O, int TASK, int NT>
__global__ void __launch_bounds__(UseSgdTeam<H>::value ? kTeamThreads : 32) op_sgd_kernel(const float *w_in, float *w_out, DataView d, float lr, int depth) {
    constexpr int P = NetSizes<I, H, O>::P;
    constexpr int IP = IPad<I>::value;
    constexpr bool TEAM = UseSgdTeam<H>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_w = reinterpret_cast<float *>(smem_raw);
    const size_t wbytes = ((size_t)P * 4 + 15) & ~(size_t)15;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + wbytes);
    float *tiles = reinterpret_cast<float *>(smem_raw + wbytes + 16);
    float *s_team = tiles + 2 * (kTileRows * IP + kTileRows);
    for (int j = threadIdx.x; j < P; j += blockDim.x) s_w[j] = w_in[j];
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
    __syncthreads();
    SgdStream st;
    st.tile_x0 = tiles; st.tile_x1 = tiles + kTileRows * IP;
    st.tile_y0 = tiles + 2 * kTileRows * IP; st.tile_y1 = tiles + 2 * kTileRows * IP + kTileRows;
    st.bar0 = &s_bar[0]; st.bar1 = &s_bar[1];
    st.parity0 = st.parity1 = 0u;
    for (int e = 0; e < depth; ++e) {                       // R:108 `depth` epochs (sgd_depth is always 1, R:170)
        if constexpr (TEAM) sgd_pass_team<I, H, O, TASK, kTeamThreads>(s_w, s_w, d, false, lr, st, s_team);
        else sgd_pass<I, H, O, TASK>(s_w, s_w, d, false, lr, st);   // always exercise the TMA-streamed path
        __syncthreads();
    }
    for (int j = threadIdx.x; j < P; j += blockDim.x) w_out[j] = s_w[j];
}

// K5 as a single operation (one CTA): evaluate_proposal / likelihood_func of a wide-hidden net on the
// tensor cores.  tiles = A tiles of the data (tc::pack_a_kernel).
template <int I, int H, int O, int TASK, int NT>
__global__ void __launch_bounds__(NT) op_forward_tc_kernel(const float *w, const float *tiles, const float *y, int n,
                                                           float *fx, float *prob, double *sums) {
    if constexpr (UseTc<I, H, O, NT>::value) {
        extern __shared__ __align__(128) unsigned char smem_raw[];
        __shared__ double s_red[3 * (NT / 32)];
        constexpr int P = NetSizes<I, H, O>::P;
        w += (size_t)blockIdx.x * P;
        if (fx) fx += (size_t)blockIdx.x * n;
        if (prob) prob += (size_t)blockIdx.x * n * O;
        sums += (size_t)blockIdx.x * 3;
        tc::State st;
        tc::setup<I, H, O>(smem_raw, st);
        tc::build_b<I, H, O>(smem_raw, w, threadIdx.x, NT);
        tc::fence_async_smem();
        __syncthreads();
        double s[3] = {0.0, 0.0, 0.0};
        int c = 0;
        tc::lik_pass<I, H, O, TASK, NT, true>(smem_raw, st, tiles, y, n, w, s[0], s[1], c, fx, prob);
        s[2] = (double)c;
        block_sum<3, NT>(s, s_red);
        if (threadIdx.x == 0) { sums[0] = s[0]; sums[1] = s[1]; sums[2] = s[2]; }
        tc::teardown<I, H, O>(st);
    }
}

}  // namespace ptfnn
