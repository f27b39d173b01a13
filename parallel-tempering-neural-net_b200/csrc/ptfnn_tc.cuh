// K5: tcgen05 likelihood pass for wide hidden layers (BASELINE configs[4], FNN 16-256-10).
//
// Network.evaluate_proposal (C:134-153) + likelihood_func (C:209-222) with a FIXED weight vector is the
// one place on the hot path where the layers really are dense GEMMs:
//     Z[N x H] = X[N x I] . W1[I x H] - B1,   hid = sigmoid(Z),   Z2[N x O] = hid[N x H] . W2[H x O] - B2.
// BOTH run on the 5th-generation tensor cores; the CUDA cores keep the sigmoids and the softmax epilogue:
//
//   * layer 1, A operand = a 128-row tile of the data set, stored ONCE per data set in HBM in the UMMA
//     canonical K-major / no-swizzle layout (8-row x 16-byte core matrices), so that one 1-D TMA
//     bulk copy (UBLKCP) lands it in shared memory ready for the tensor core.  The tile is augmented
//     with a constant-one column: the bias is part of the GEMM.  B operand = -log2(e) W1^T (and +log2(e) B1
//     in the bias row), rebuilt per proposal in shared memory: the accumulator is born in the ex2 domain
//     of the sigmoid.  tcgen05.mma.cta_group::1.kind::tf32 (SS), M = 128, N = 128 (one half of the hidden
//     layer at a time), K = 8 per instruction.
//   * layer 2, A operand = the hidden activations IN TENSOR MEMORY: the epilogue threads read 32 columns of
//     pre-activations (tcgen05.ld), apply the grouped-reciprocal sigmoid and write the activations back IN PLACE
//     (tcgen05.st) -- the accumulator columns of layer 1 become the A operand of layer 2 (TS-mode MMA,
//     A[row = lane][k = column]).  B operand = -log2(e) W2^T padded to 16 outputs, in shared memory.  Round 1
//     evaluated this layer on the CUDA cores: 2560 FMAs and 640 broadcast LDS.128 per row kept the FMA pipe
//     and the shared-memory port busier than the XU pipe the sigmoids need (tensor pipe 12 %, shared-memory
//     wavefronts 59 %).
//   * fp32 accuracy from the tf32 pipe: every operand is split into tf32 "hi" and "lo" parts and
//     D = A_hi B_hi + A_lo B_hi + A_hi B_lo (3xTF32, error ~2^-21).  Plain tf32 (2^-11) would put
//     ~1e-2 of noise on a 20 000-row log-likelihood, i.e. on the Metropolis-Hastings decision.  For layer 2
//     the hi / lo parts of W2 sit side by side in one N = 32 operand (one MMA gives A_hi B_hi | A_hi B_lo),
//     the lo part of the activations is a second, N = 16 MMA: 2 instead of 3 instructions per K step --
//     a CTA issues one small MMA per ~56 cycles whatever its N (tools/ts_mma_probe.cu), so their NUMBER is
//     what layer 2 costs.
//   * TMEM, 256 columns per CTA (two temperatures per SM): Z = 128 (layer-1 accumulator of one half of the hidden
//     layer, overwritten by hid_hi), L = 2 x 32 (hid_lo of the sub-block in flight, double buffered), D2 = 32
//     (layer-2 accumulator: 16 columns hi.hi + lo.hi, 16 columns hi.lo).  MMAs of one CTA execute in issue order,
//     so layer 1 of the next half is issued right behind the layer-2 MMAs that still read Z.
//   * warp-specialised: four epilogue warps (one per TMEM lane quadrant, thread = row) and one MMA warp whose elected
//     lane issues every MMA; hand-offs through mbarriers only (tcgen05.commit -> mbarrier towards the epilogue, one
//     arrival per epilogue warp towards the MMA warp) -- no CTA-wide barrier inside the pass.
//
// The XU pipe (256 sigmoids per row at 16 lanes / clk / SM) bounds the pass, as SURVEY 7 predicted.
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
// the same load without the wait: the registers may be read after tmem_wait_ld() only
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
// (the registers are in/out operands of the wait so that no use of them can be scheduled above it)
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// shared-memory plan of the pass (bytes; all parts 128-byte aligned)
// ------------------------------------------------------------------------------------------
constexpr int kN2 = 16;                          // outputs padded to the smallest N of an M = 128 MMA
constexpr int kSub = 32;                         // hidden units per epilogue sub-block (one tcgen05.ld / st)
constexpr int kHalf = 128;                       // hidden units per layer-1 MMA (N)
constexpr int kTmemCols = 256;                   // Z 128 | L 2 x 32 | D2 32 | spare 32
constexpr int kColZ = 0, kColL = 128, kColD = 192;

template <int I, int H, int O>
struct Smem {
    static constexpr int off_b = 0;                                          // layer 1: hi | lo, CH chunks x H rows x 16 B
    static constexpr int off_b2 = off_b + BGeometry<I, H>::B_BYTES;           // layer 2: H/4 chunks x (16 hi rows + 16 lo rows) x 16 B
    static constexpr int B2_CHUNK_BYTES = 2 * kN2 * 16;
    static constexpr int B2_BYTES = (H / 4) * B2_CHUNK_BYTES;
    static constexpr int off_a = off_b2 + B2_BYTES;
    static constexpr int off_bar = off_a + Geometry<I>::A_TILE_BYTES;        // a_full, z_full, l_free[2], d_full, tmem slot
    static constexpr int total = off_bar + 128;
};

// Both B operands of a weight vector w (layout a1).  Layer 1: row n (hidden unit n) of every chunk.  Layer 2: chunk
// c holds hidden units 4c .. 4c+3 (K) for the 16 (padded) outputs: rows 0-15 the tf32 hi part, rows 16-31 the lo part.
template <int I, int H, int O>
__device__ __forceinline__ void build_b(unsigned char *smem, const float *w, int tid, int nt) {
    using G = Geometry<I>;
    using BG = BGeometry<I, H>;
    using S = Smem<I, H, O>;
    constexpr int oW2 = I * H, oB1 = I * H + H * O;
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
            *reinterpret_cast<float4 *>(smem + S::off_b + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(smem + S::off_b + BG::B_HALF_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    for (int idx = tid; idx < (H / 4) * kN2; idx += nt) {
        const int c = idx / kN2, o = idx % kN2;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float v = o < O ? -kL2E * w[oW2 + (c * 4 + e) * O + o] : 0.0f;     // -log2(e) W2[h][o]: z2 in the ex2 domain too
            hi[e] = tf32_hi(v);
            lo[e] = tf32_hi(v - hi[e]);
        }
        unsigned char *row = smem + S::off_b2 + c * S::B2_CHUNK_BYTES + o * 16;
        *reinterpret_cast<float4 *>(row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(row + kN2 * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// TS-mode MMA: A from tensor memory (rows = lanes, K = consecutive columns), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// 32 registers per thread -> 32 lanes x 32 consecutive columns of TMEM (lane = TMEM lane)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
          "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
          "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
          "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#ifdef PTFNN_TC_TRACE      // measurement builds only (tools/tc_trace.cu): clock stamps of the first tiles, per role
__device__ long long g_tc_trace[2][8][64];
#define TC_STAMP(role, k) do { if (blockIdx.x == 0 && lane == 0 && (role == 1 || warp == 0) && t < 8) g_tc_trace[role][t][k] = clock64(); } while (0)
#else
#define TC_STAMP(role, k) do { } while (0)
#endif
constexpr int kEpiThreads = 128;                 // warps 0-3: one per TMEM lane quadrant (thread = row of the tile)
constexpr int kThreads = 160;                    // + warp 4: issues every MMA (a CTA issues one small MMA per ~56 cycles;
                                                 //   done by an epilogue warp, that time was added to the sigmoids' instead of hidden under them)
// mbarriers of the pass
enum { kBarA = 0, kBarZ = 1, kBarL0 = 2, kBarL1 = 3, kBarD = 4, kBarH0 = 5, kBarH1 = 6, kNumBars = 7 };

struct State {
    uint32_t tmem;                      // TMEM base address of the CTA's 256 columns
    // mbarrier phases, each tracked by the role that waits on it
    uint32_t ph_a, ph_h0, ph_h1;        // issuer: A tile landed; hid of sub-block buffer b stored by all four warps
    uint32_t ph_z, ph_d, ph_l0, ph_l1;  // epilogue: Z full; D2 full; L[b] consumed
    bool pend_l0, pend_l1;              // epilogue: a commit on L[b] that has not been waited for yet
};

// One-time set-up by the whole CTA: mbarriers, TMEM allocation (warp 0).
template <int I, int H, int O>
__device__ __forceinline__ void setup(unsigned char *smem, State &st) {
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Smem<I, H, O>::off_bar);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bars + kNumBars);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kNumBars; ++k) mbar_init(&bars[k], (k == kBarH0 || k == kBarH1) ? kEpiThreads / 32 : 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    st.tmem = *slot;
    st.ph_a = st.ph_h0 = st.ph_h1 = 0u;
    st.ph_z = st.ph_d = st.ph_l0 = st.ph_l1 = 0u;
    st.pend_l0 = st.pend_l1 = false;
}
template <int I, int H, int O>
__device__ __forceinline__ void teardown(State &st) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(st.tmem, kTmemCols);
}

// The pass over one data set.  tiles = A tiles of the data set (pack_a_kernel), y = labels, n rows.
// Both B operands must have been built (build_b) and made visible (fence_async_smem + __syncthreads) by the caller.
// w points at the SAME weight vector (layout a1, anywhere readable): only B2 is read from it here.
// Called by the whole CTA (kThreads); the caller synchronises the CTA afterwards (block_sum does).
//
//   issuer (warp 4, one lane)                         epilogue (warps 0-3, thread = row)
//   wait A tile; layer 1 (half 0) -> Z; commit Z      wait Z
//   per 32-unit sub-block g, buffer b = g & 1:        ld z; sigmoid; hi -> Z in place, lo -> L[b] (after L[b] is free);
//     wait H[b]                                         fence; arrive H[b]  (one lane per warp)
//     layer 2 from Z / L[b] -> D2
//     commit L[b]  |  + layer 1 of the next half, commit Z  |  commit D (last sub-block of the tile)
//                                                      wait D; ld D2; output sigmoids, softmax, sums
// MMAs execute in issue order, so layer 1 of the next half / tile is issued right behind the layer-2 MMAs that still read
// Z; the epilogue's own reads of Z and D2 precede its arrival on H, which precedes every MMA that overwrites them.
template <int I, int H, int O, int TASK, int NT, bool WRITE>
__device__ __forceinline__ void lik_pass(unsigned char *smem, State &st, const float *__restrict__ tiles,
                                         const float *__restrict__ y, int n, const float *__restrict__ w,
                                         double &s0, double &s1, int &correct, float *fx_out, float *prob_out) {
    using G = Geometry<I>;
    using BG = BGeometry<I, H>;
    using S = Smem<I, H, O>;
    static_assert(NT == kThreads && H % kHalf == 0 && H <= 256 && O <= kN2, "four epilogue warps (thread = row) + the MMA warp");
    constexpr int oB2 = I * H + H * O + H;
    constexpr uint32_t IDESC1 = instr_desc_tf32(kRows, kHalf);
    constexpr uint32_t IDESC2A = instr_desc_tf32(kRows, 2 * kN2), IDESC2B = instr_desc_tf32(kRows, kN2);
    constexpr int NHALF = H / kHalf, NSUB = kHalf / kSub;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::off_bar);
    const uint32_t a_base = smem_u32(smem + S::off_a), b_base = smem_u32(smem + S::off_b), b2_base = smem_u32(smem + S::off_b2);
    const uint32_t tZ = st.tmem + kColZ, tL = st.tmem + kColL, tD = st.tmem + kColD;
    const int ntiles = (n + kRows - 1) / kRows;
    if (ntiles <= 0) return;

    if (warp == kEpiThreads / 32) {
        // ================= the MMA warp =================
        if (lane == 0) {
            // layer 1 of one half of the hidden layer: Z = A_hi B_hi + A_lo B_hi + A_hi B_lo (N = 128 rows of B from `half`)
            auto issue_layer1 = [&](int half) {
                const uint32_t bh = b_base + (uint32_t)half * kHalf * 16;
                bool acc = false;
#pragma unroll
                for (int ks = 0; ks < G::CH / 2; ++ks) {
                    mma_tf32(tZ, smem_desc(a_base + ks * 2 * G::A_CHUNK_BYTES, G::A_CHUNK_BYTES, 128),
                             smem_desc(bh + ks * 2 * BG::B_CHUNK_BYTES, BG::B_CHUNK_BYTES, 128), IDESC1, acc);
                    acc = true;
                }
#pragma unroll
                for (int ks = 0; ks < G::CL / 2; ++ks)
                    mma_tf32(tZ, smem_desc(a_base + (G::CH + ks * 2) * G::A_CHUNK_BYTES, G::A_CHUNK_BYTES, 128),
                             smem_desc(bh + ks * 2 * BG::B_CHUNK_BYTES, BG::B_CHUNK_BYTES, 128), IDESC1, true);
#pragma unroll
                for (int ks = 0; ks < G::CH / 2; ++ks)
                    mma_tf32(tZ, smem_desc(a_base + ks * 2 * G::A_CHUNK_BYTES, G::A_CHUNK_BYTES, 128),
                             smem_desc(bh + BG::B_HALF_BYTES + ks * 2 * BG::B_CHUNK_BYTES, BG::B_CHUNK_BYTES, 128), IDESC1, true);
            };
            for (int t = 0; t < ntiles; ++t) {
                TC_STAMP(1, 0);
                mbar_wait_suspended(&bars[kBarA], st.ph_a);
                st.ph_a ^= 1u;
                tc_fence_after();
                TC_STAMP(1, 1);
                issue_layer1(0);
                mma_commit(&bars[kBarZ]);
                TC_STAMP(1, 2);
#pragma unroll
                for (int half = 0; half < NHALF; ++half) {
#pragma unroll
                    for (int sb = 0; sb < NSUB; ++sb) {
                        const int b = sb & 1;
                        if (b) { mbar_wait_suspended(&bars[kBarH1], st.ph_h1); st.ph_h1 ^= 1u; }
                        else { mbar_wait_suspended(&bars[kBarH0], st.ph_h0); st.ph_h0 ^= 1u; }
                        tc_fence_after();
                        TC_STAMP(1, 3 + 2 * (half * NSUB + sb));
                        // layer 2, K = 32 hidden units of this sub-block: hid_hi . [W2_hi | W2_lo] (N = 32), hid_lo . W2_hi (N = 16)
#pragma unroll
                        for (int ks = 0; ks < kSub / 8; ++ks) {
                            const uint32_t bk = b2_base + (uint32_t)(((half * kHalf + sb * kSub) / 4 + ks * 2) * S::B2_CHUNK_BYTES);
                            const uint64_t bd = smem_desc(bk, S::B2_CHUNK_BYTES, 128);
                            mma_tf32_ts(tD, tZ + (uint32_t)(sb * kSub + ks * 8), bd, IDESC2A, !(half == 0 && sb == 0 && ks == 0));
                            mma_tf32_ts(tD, tL + (uint32_t)(b * kSub + ks * 8), bd, IDESC2B, true);
                        }
                        if (sb < NSUB - 1) {
                            mma_commit(&bars[b ? kBarL1 : kBarL0]);
                        } else if (half + 1 < NHALF) {
                            issue_layer1(half + 1);            // executes behind the MMAs above (issue order): Z is free by then
                            mma_commit(&bars[kBarZ]);
                        } else {
                            mma_commit(&bars[kBarD]);
                        }
                        TC_STAMP(1, 4 + 2 * (half * NSUB + sb));
                    }
                }
            }
        }
        return;
    }

    // ================= the epilogue warps =================
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;          // this warp's TMEM lane quadrant
    float b2l[O];
#pragma unroll
    for (int o = 0; o < O; ++o) b2l[o] = kL2E * w[oB2 + o];
    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[kBarA], G::A_TILE_BYTES);
        tma_load_1d(smem + S::off_a, tiles, G::A_TILE_BYTES, &bars[kBarA]);
    }
    for (int t = 0; t < ntiles; ++t) {
        // this row's label: requested now, used after the whole tile (a global load in the tile epilogue cost ~600 cycles)
        const int row_in = warp * 32 + lane;
        const int r = t * kRows + row_in;
        const float yv = y[r < n ? r : n - 1];
#pragma unroll
        for (int half = 0; half < NHALF; ++half) {
            TC_STAMP(0, 40 + 2 * half);
            mbar_wait(&bars[kBarZ], st.ph_z);          // layer-1 pre-activations of this half are in Z (and every earlier MMA is done)
            st.ph_z ^= 1u;
            tc_fence_after();
            TC_STAMP(0, 41 + 2 * half);
            if (half == NHALF - 1 && tid == 0 && t + 1 < ntiles) {     // layer 1 has consumed the A tile: fetch the next one under the epilogue
                mbar_arrive_expect_tx(&bars[kBarA], G::A_TILE_BYTES);
                tma_load_1d(smem + S::off_a, tiles + (size_t)(t + 1) * G::A_TILE_FLOATS, G::A_TILE_BYTES, &bars[kBarA]);
            }
            // the pre-activations of sub-block sb + 1 are fetched from tensor memory while sub-block sb is computed
            // (two register buffers, the loop unrolled so that they alternate without copies)
            uint32_t zr[2][32];
            tmem_ld32_issue(tZ + lane_base, zr[0]);
#pragma unroll
            for (int sb = 0; sb < NSUB; ++sb) {
                const int b = sb & 1;
                tmem_wait_ld(zr[b]);
                if (sb + 1 < NSUB) tmem_ld32_issue(tZ + lane_base + (uint32_t)((sb + 1) * kSub), zr[b ^ 1]);
                TC_STAMP(0, 4 * (half * NSUB + sb));
                uint32_t lo[32];               // (the hi parts go back into zr[b]: the same registers they were loaded into)
#pragma unroll
                for (int g = 0; g < 32; g += 4) {
                    f2_t zz[2] = {pack2(__uint_as_float(zr[b][g]), __uint_as_float(zr[b][g + 1])),
                                  pack2(__uint_as_float(zr[b][g + 2]), __uint_as_float(zr[b][g + 3]))}, hh[2];
                    sigmoid_pairs<2>(zz, hh);
                    float h4[4];
                    unpack2(hh[0], h4[0], h4[1]);
                    unpack2(hh[1], h4[2], h4[3]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        // hi = the leading 11 bits (truncated: one LOP3; cvt.rna.tf32 is a four-instruction sequence on sm_100),
                        // lo = the exact rest, of which the tensor core reads the leading 11 bits: 2^-21 relative to hid in all
                        const uint32_t hb = __float_as_uint(h4[q]) & 0xffffe000u;
                        zr[b][g + q] = hb;
                        lo[g + q] = __float_as_uint(h4[q] - __uint_as_float(hb));
                    }
                }
                TC_STAMP(0, 4 * (half * NSUB + sb) + 1);
                if (b ? st.pend_l1 : st.pend_l0) {             // the layer-2 MMAs that read L[b] two sub-blocks ago
                    mbar_wait(&bars[b ? kBarL1 : kBarL0], b ? st.ph_l1 : st.ph_l0);
                    if (b) { st.ph_l1 ^= 1u; st.pend_l1 = false; } else { st.ph_l0 ^= 1u; st.pend_l0 = false; }
                    tc_fence_after();
                }
                TC_STAMP(0, 4 * (half * NSUB + sb) + 2);
                tmem_st32(tZ + lane_base + (uint32_t)(sb * kSub), zr[b]);
                tmem_st32(tL + lane_base + (uint32_t)(b * kSub), lo);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[b ? kBarH1 : kBarH0]);      // this warp's 32 rows of the sub-block are in tensor memory
                TC_STAMP(0, 4 * (half * NSUB + sb) + 3);
                if (sb < NSUB - 1) { if (b) st.pend_l1 = true; else st.pend_l0 = true; }
            }
        }
        // ---- output layer of the tile: D2 = hi.hi + lo.hi | hi.lo
        TC_STAMP(0, 44);
        mbar_wait(&bars[kBarD], st.ph_d);
        st.ph_d ^= 1u;
        tc_fence_after();
        TC_STAMP(0, 45);
        float d2[32];
        tmem_ld32(tD + lane_base, d2);
        tc_fence_before();                                     // (ordered before this warp's next arrival on H: D2 is rewritten after it)
        if (r < n) {
            float out[O];
#pragma unroll
            for (int o = 0; o < O; ++o) out[o] = rcp_ftz(1.0f + ex2_ftz(d2[o] + d2[kN2 + o] + b2l[o]));     // R:54-55
            if constexpr (TASK == kTaskReg) {
                const float e = yv - out[0];
                s0 += (double)(e * e);
                if constexpr (WRITE) fx_out[r] = out[0];
            } else {
                int am = 0;
                float se = 0.0f, om = out[0];
#pragma unroll
                for (int o = 0; o < O; ++o) se += __expf(out[o]);                // C:108-110 (arguments in (0,1): ex2.approx is good to ~2 ulp)
#pragma unroll
                for (int o = 1; o < O; ++o) { const bool g = out[o] > om; am = g ? o : am; om = g ? out[o] : om; }   // np.argmax: first max
                const int lab = (int)yv;
                float ol = out[0];
#pragma unroll
                for (int o = 1; o < O; ++o) ol = (o == lab) ? out[o] : ol;
                s0 += (double)(ol - __logf(se));
                const float e = (float)am - yv;
                s1 += (double)(e * e);
                correct += ((float)am == yv) ? 1 : 0;
                if constexpr (WRITE) {
                    fx_out[r] = (float)am;
                    if (prob_out) {
#pragma unroll
                        for (int o = 0; o < O; ++o) prob_out[(size_t)r * O + o] = __expf(out[o]) / se;
                    }
                }
            }
        }
    }
}

}  // namespace tc
}  // namespace ptfnn
