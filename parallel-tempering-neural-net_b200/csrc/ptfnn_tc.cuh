// K5: tcgen05 likelihood pass for wide hidden layers (BASELINE configs[4], FNN 16-256-10).
//
// Network.evaluate_proposal (C:134-153) + likelihood_func (C:209-222) with a FIXED weight vector is the
// one place on the hot path where a layer really is a dense GEMM:  Z[N x H] = X[N x I] . W1[I x H] - B1.
// It runs on the 5th-generation tensor cores:
//
//   * A operand = a 128-row tile of the data set, stored ONCE per data set in HBM in the UMMA
//     canonical K-major / no-swizzle layout (8-row x 16-byte core matrices), so that one 1-D TMA
//     bulk copy (UBLKCP) lands it in shared memory ready for the tensor core.  The tile is augmented
//     with a constant-one column: the bias is part of the GEMM.
//   * B operand = -log2(e) W1^T (and +log2(e) B1 in the bias row), rebuilt per proposal in shared
//     memory: the accumulator is born in the ex2 domain of the sigmoid.
//   * fp32 accuracy from the tf32 pipe: both operands are split into tf32 "hi" and "lo" parts and
//     D = A_hi B_hi + A_lo B_hi + A_hi B_lo (3xTF32, error ~2^-21).  Plain tf32 (2^-11) would put
//     ~1e-2 of noise on a 20 000-row log-likelihood, i.e. on the Metropolis-Hastings decision.
//   * tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = H = 256, K = 8 per instruction, issued by one
//     elected thread; accumulator 128 lanes x 256 columns of TMEM (two temperatures per SM = all 512
//     columns); completion through tcgen05.commit -> mbarrier.
//   * epilogue on the CUDA cores straight from TMEM (tcgen05.ld 32x32b.x32): grouped-reciprocal
//     sigmoids, the H x O output layer with packed FFMA2, softmax-of-sigmoid log-likelihood (C:108-110,
//     C:215-219).  Each lane quadrant of TMEM is read by two warps (128 hidden units each); the two
//     halves of a row meet through shared memory.
//
// The epilogue, not the tensor pipe, bounds the pass (256 sigmoids per row on the 16-lane MUFU unit),
// as SURVEY 7 predicted; the tensor cores take the 2 I H flop of layer 1 off the FMA pipe.
#pragma once
#include "ptfnn_device.cuh"

namespace ptfnn {
namespace tc {

constexpr int kRows = 128;                       // rows per tile = UMMA M

template <int I>
struct Geometry {
    static constexpr int XC = (I + 3) / 4;                       // 16-byte K chunks holding x
    static constexpr int CH = 2 * ((XC + 1 + 1) / 2);            // hi chunks: x, the bias chunk, zero pad to a K = 8 step
    static constexpr int CL = 2 * ((XC + 1) / 2);                // lo chunks: x, zero pad
    static constexpr int A_CHUNK_BYTES = kRows * 16;
    static constexpr int A_TILE_BYTES = (CH + CL) * A_CHUNK_BYTES;
    static constexpr int A_TILE_FLOATS = A_TILE_BYTES / 4;
};
template <int I, int H>
struct BGeometry {
    static constexpr int CH = Geometry<I>::CH;
    static constexpr int B_CHUNK_BYTES = H * 16;
    static constexpr int B_HALF_BYTES = CH * B_CHUNK_BYTES;      // hi (or lo) part
    static constexpr int B_BYTES = 2 * B_HALF_BYTES;
};
__host__ __device__ constexpr int a_tile_floats(int I) {
    return ((2 * (((I + 3) / 4 + 2) / 2)) + (2 * (((I + 3) / 4 + 1) / 2))) * kRows * 4;
}

__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// Data set -> A tiles.  One thread per (tile row, chunk).  x is the padded row-major [n][IP] array.
// tile layout: [hi chunk 0 .. CH) [lo chunk 0 .. CL), chunk = 128 rows x 16 bytes.
template <int I>
__global__ void pack_a_kernel(const float *__restrict__ x, int n, int IP, float *__restrict__ tiles) {
    using G = Geometry<I>;
    const int ntiles = (n + kRows - 1) / kRows;
    const long long total = (long long)ntiles * kRows * (G::CH + G::CL);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int row_in = (int)(idx % kRows);
        const int c = (int)((idx / kRows) % (G::CH + G::CL));
        const int t = (int)(idx / ((long long)kRows * (G::CH + G::CL)));
        const int r = t * kRows + row_in;
        const bool lo = c >= G::CH;
        const int ck = lo ? c - G::CH : c;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = ck * 4 + e;
                if (k < I) {
                    const float xv = x[(size_t)r * IP + k];
                    const float h = tf32_hi(xv);
                    v[e] = lo ? tf32_hi(xv - h) : h;
                } else if (k == G::XC * 4 && !lo) {
                    v[e] = 1.0f;                                   // the bias column
                }
            }
        }
        reinterpret_cast<float4 *>(tiles)[(size_t)t * (G::A_TILE_FLOATS / 4) + (size_t)c * kRows + row_in] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor):
// core matrix = 8 rows x 16 bytes, rows 16 bytes apart; SBO = distance between 8-row groups,
// LBO = distance between the two 16-byte K chunks of one K = 8 step.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    return d;                                      // base offset 0, LBO mode 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive columns of TMEM -> 32 registers per thread (lane = TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------
// shared-memory plan of the pass (bytes; all parts 128-byte aligned)
// ------------------------------------------------------------------------------------------
template <int I, int H, int O>
struct Smem {
    static constexpr int off_b = 0;
    static constexpr int off_a = off_b + BGeometry<I, H>::B_BYTES;
    static constexpr int off_x = off_a + Geometry<I>::A_TILE_BYTES;          // [kRows][XO] partial output sums of the upper half
    static constexpr int XO = (O + 3) & ~3;
    static constexpr int off_bar = off_x + kRows * XO * 4;                   // a_full, mma_done, tmem slot
    static constexpr int total = off_bar + 32;
};
__host__ __device__ constexpr int tc_smem_bytes(int I, int H, int O) {
    return 2 * (2 * (((I + 3) / 4 + 2) / 2)) * H * 16 + a_tile_floats(I) * 4 + kRows * ((O + 3) & ~3) * 4 + 32;
}

// B operand of a weight vector w (layout a1): thread n < H writes row n (hidden unit n) of every chunk.
template <int I, int H, int O>
__device__ __forceinline__ void build_b(unsigned char *smem, const float *w, int tid, int nt) {
    using G = Geometry<I>;
    using BG = BGeometry<I, H>;
    constexpr int oB1 = I * H + H * O;
    for (int n = tid; n < H; n += nt) {
#pragma unroll
        for (int c = 0; c < G::CH; ++c) {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = c * 4 + e;
                float v = 0.0f;
                if (k < I) v = -kL2E * w[k * H + n];                  // -log2(e) W1[k][n]
                else if (k == G::XC * 4) v = kL2E * w[oB1 + n];       // bias row: z = x.W1 - B1 (R:52) in the ex2 domain
                hi[e] = tf32_hi(v);
                lo[e] = tf32_hi(v - hi[e]);
            }
            const int off = c * BG::B_CHUNK_BYTES + n * 16;
            *reinterpret_cast<float4 *>(smem + Smem<I, H, O>::off_b + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(smem + Smem<I, H, O>::off_b + BG::B_HALF_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

struct State {
    uint32_t tmem;          // TMEM base address of the 128 x H accumulator
    uint32_t ph_a, ph_mma;  // mbarrier phases (every thread tracks them)
};

// One-time set-up by the whole CTA: mbarriers, TMEM allocation (warp 0).
template <int I, int H, int O>
__device__ __forceinline__ void setup(unsigned char *smem, State &st) {
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Smem<I, H, O>::off_bar);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 2);
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(slot, H);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    st.tmem = *slot;
    st.ph_a = st.ph_mma = 0u;
}
template <int I, int H, int O>
__device__ __forceinline__ void teardown(State &st) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(st.tmem, H);
}

// The pass over one data set.  tiles = A tiles of the data set (pack_a_kernel), y = labels, n rows.
// B must have been built (build_b) and made visible (fence_async_smem + __syncthreads) by the caller.
// w2 / b2 point at the output layer of the SAME weight vector (layout a1, anywhere readable).
template <int I, int H, int O, int TASK, int NT, bool WRITE>
__device__ __forceinline__ void lik_pass(unsigned char *smem, State &st, const float *__restrict__ tiles,
                                         const float *__restrict__ y, int n, const float *__restrict__ w,
                                         double &s0, double &s1, int &correct, float *fx_out, float *prob_out) {
    using G = Geometry<I>;
    using BG = BGeometry<I, H>;
    using S = Smem<I, H, O>;
    static_assert(NT == 128 && H % 32 == 0 && H <= 256, "one warp per TMEM lane quadrant: thread = row, all H accumulator columns");
    static_assert(O % 2 == 0 && (H * O) % 4 == 0, "output-layer rows are read in aligned pairs");
    constexpr int oW2 = I * H, oB2 = I * H + H * O + H;
    constexpr uint32_t IDESC = instr_desc_tf32(kRows, H);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int quad = warp & 3;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::off_bar);
    uint64_t *bar_a = &bars[0], *bar_mma = &bars[1];
    const uint32_t a_base = smem_u32(smem + S::off_a), b_base = smem_u32(smem + S::off_b);
    const int ntiles = (n + kRows - 1) / kRows;

    if (tid == 0 && ntiles > 0) {
        mbar_arrive_expect_tx(bar_a, G::A_TILE_BYTES);
        tma_load_1d(smem + S::off_a, tiles, G::A_TILE_BYTES, bar_a);
    }
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0) {
            mbar_wait(bar_a, st.ph_a);
            tc_fence_after();
            // D = A_hi B_hi + A_lo B_hi + A_hi B_lo, K = 8 per instruction (two 16-byte chunks)
            bool acc = false;
#pragma unroll
            for (int ks = 0; ks < G::CH / 2; ++ks) {
                mma_tf32(st.tmem, smem_desc(a_base + ks * 2 * G::A_CHUNK_BYTES, G::A_CHUNK_BYTES, 128),
                         smem_desc(b_base + ks * 2 * BG::B_CHUNK_BYTES, BG::B_CHUNK_BYTES, 128), IDESC, acc);
                acc = true;
            }
#pragma unroll
            for (int ks = 0; ks < G::CL / 2; ++ks)
                mma_tf32(st.tmem, smem_desc(a_base + (G::CH + ks * 2) * G::A_CHUNK_BYTES, G::A_CHUNK_BYTES, 128),
                         smem_desc(b_base + ks * 2 * BG::B_CHUNK_BYTES, BG::B_CHUNK_BYTES, 128), IDESC, true);
#pragma unroll
            for (int ks = 0; ks < G::CH / 2; ++ks)
                mma_tf32(st.tmem, smem_desc(a_base + ks * 2 * G::A_CHUNK_BYTES, G::A_CHUNK_BYTES, 128),
                         smem_desc(b_base + BG::B_HALF_BYTES + ks * 2 * BG::B_CHUNK_BYTES, BG::B_CHUNK_BYTES, 128), IDESC, true);
            mma_commit(bar_mma);
        }
        st.ph_a ^= 1u;
        mbar_wait(bar_mma, st.ph_mma);
        st.ph_mma ^= 1u;
        tc_fence_after();
        if (tid == 0 && t + 1 < ntiles) {           // the MMAs have consumed the A tile: fetch the next one under the epilogue
            mbar_arrive_expect_tx(bar_a, G::A_TILE_BYTES);
            tma_load_1d(smem + S::off_a, tiles + (size_t)(t + 1) * G::A_TILE_FLOATS, G::A_TILE_BYTES, bar_a);
        }
        // ---- epilogue: this thread = row (32 quad + lane) of the tile, all H hidden units
        f2_t acc2[O / 2];
#pragma unroll
        for (int j = 0; j < O / 2; ++j) acc2[j] = pack2(0.0f, 0.0f);
#pragma unroll 1
        for (int cb = 0; cb < H / 32; ++cb) {
            float z[32];
            const int h0 = cb * 32;
            tmem_ld32(st.tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)h0, z);
#pragma unroll
            for (int g = 0; g < 32; g += 4) {
                f2_t zz[2] = {pack2(z[g], z[g + 1]), pack2(z[g + 2], z[g + 3])}, hh[2];
                sigmoid_pairs<2>(zz, hh);
                float hid[4];
                unpack2(hh[0], hid[0], hid[1]);
                unpack2(hh[1], hid[2], hid[3]);
#pragma unroll
                for (int q = 0; q < 4; q += 2) {
                    // output-layer rows of hidden units h, h+1: 2 O contiguous floats (layout a1), 16-byte aligned
                    const float4 *wr = reinterpret_cast<const float4 *>(w + oW2 + (h0 + g + q) * O);
                    float wv[2 * O];
#pragma unroll
                    for (int v4 = 0; v4 < (2 * O) / 4; ++v4) {
                        const float4 u = wr[v4];
                        wv[4 * v4] = u.x; wv[4 * v4 + 1] = u.y; wv[4 * v4 + 2] = u.z; wv[4 * v4 + 3] = u.w;
                    }
#pragma unroll
                    for (int j = 0; j < O / 2; ++j) {
                        acc2[j] = fma2(pack2(wv[2 * j], wv[2 * j + 1]), pack2(hid[q], hid[q]), acc2[j]);
                        acc2[j] = fma2(pack2(wv[O + 2 * j], wv[O + 2 * j + 1]), pack2(hid[q + 1], hid[q + 1]), acc2[j]);
                    }
                }
            }
        }
        float part[O];
#pragma unroll
        for (int j = 0; j < O / 2; ++j) unpack2(acc2[j], part[2 * j], part[2 * j + 1]);
        const int row_in = quad * 32 + lane;
        tc_fence_before();
        __syncthreads();                              // every TMEM read of this tile is done: the next tile's MMAs may overwrite it
        const int r = t * kRows + row_in;
        if (r < n) {
            float out[O];
            const float yv = y[r];
#pragma unroll
            for (int o = 0; o < O; ++o) out[o] = sigmoid_fast(part[o] - w[oB2 + o]);   // R:54-55
            if constexpr (TASK == kTaskReg) {
                const float e = yv - out[0];
                s0 += (double)(e * e);
                if constexpr (WRITE) fx_out[r] = out[0];
            } else {
                int am = 0;
                float se = 0.0f;
#pragma unroll
                for (int o = 0; o < O; ++o) se += expf(out[o]);                  // C:108-110
#pragma unroll
                for (int o = 1; o < O; ++o) am = (out[o] > out[am]) ? o : am;    // np.argmax: first max
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
        tc_fence_after();
    }
}

}  // namespace tc
}  // namespace ptfnn
