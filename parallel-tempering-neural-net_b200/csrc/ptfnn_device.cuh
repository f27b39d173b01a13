// Device-side building blocks of the sm_100a parallel-tempering FNN sampler.
// Nothing here is derived from reference code (the reference has no native code at all);
// file:line citations point at the reference behaviour each piece has to reproduce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptfnn {

constexpr int kTaskReg = 0;
constexpr int kTaskCls = 1;
constexpr int kWarp = 32;

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (free-running mode).  __host__ __device__ so that the
// host can reproduce the integer stream bit-for-bit (swap uniforms on the multi-GPU path).
// Counter layout: {step | round, block, stream, tag}; key = 64-bit seed.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kTagProposal = 0x50524f50u;  // lx, eta noise, proposal normals  (R:327-355)
constexpr uint32_t kTagAccept = 0x41434350u;    // MH uniform                        (R:387)
constexpr uint32_t kTagSwap = 0x53574150u;      // coordinator swap uniforms         (R:677)
constexpr uint32_t kStreamCommon = 0xffffffffu; // common-random-numbers stream      

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__host__ __device__ __forceinline__ void philox_draw(uint64_t seed, uint32_t a, uint32_t b, uint32_t s,
                                                     uint32_t tag, uint32_t (&out)[4]) {
    out[0] = a; out[1] = b; out[2] = s; out[3] = tag;
    philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 24-bit uniforms: exactly representable in float32, so host and device agree bit-for-bit.
__host__ __device__ __forceinline__ float u01_open_right(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }      // [0,1)
__host__ __device__ __forceinline__ float u01_open_left(uint32_t x) { return (float)((x >> 8) + 1u) * (1.0f / 16777216.0f); } // (0,1]

#ifdef __CUDACC__
// Box-Muller; two uniforms -> two standard normals.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &z0, float &z1) {
    const float r = sqrtf(-2.0f * logf(u01_open_left(a)));
    float s, c;
    sincospif(2.0f * u01_open_right(b), &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// The draws of one replica-step, shared by the chain kernel and ptfnn_generate_draws so that the
// verification dump is the same code path.
struct StepDraws {
    float lx, z_eta, u;
};
__device__ __forceinline__ StepDraws philox_step_scalars(uint64_t seed, uint32_t step, uint32_t stream,
                                                         uint32_t replica) {
    uint32_t c[4];
    StepDraws d;
    philox_draw(seed, step, 0u, stream, kTagProposal, c);
    d.lx = u01_open_right(c[0]);
    float z1;
    box_muller(c[2], c[3], d.z_eta, z1);
    philox_draw(seed, step, 0u, replica, kTagAccept, c);
    d.u = u01_open_right(c[0]);
    return d;
}
// normals z[4*blk .. 4*blk+3] of the proposal noise vector
__device__ __forceinline__ void philox_step_normals4(uint64_t seed, uint32_t step, uint32_t stream, uint32_t blk,
                                                     float (&z)[4]) {
    uint32_t c[4];
    philox_draw(seed, step, 1u + blk, stream, kTagProposal, c);
    box_muller(c[0], c[1], z[0], z[1]);
    box_muller(c[2], c[3], z[2], z[3]);
}

// ------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------
// Latency-critical sigmoid of the serial SGD recurrence: MUFU.EX2 + MUFU.RCP (R:43-44).
// ex2.approx.ftz / rcp.approx.ftz directly: no denormal-range fix-up code on the dependent chain
// (ftz only flushes results below 2^-126, where the sigmoid is 1 or 0 to fp32 precision anyway).
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float z) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * z)); }
// Sigmoid of the row-parallel likelihood pass: full-precision expf and IEEE division.
__device__ __forceinline__ float sigmoid_precise(float z) { return 1.0f / (1.0f + expf(-z)); }

// RB sigmoids with ONE reciprocal.  The row-parallel likelihood pass is bound by the XU (MUFU) pipe
// (16 lanes/clk/SM on B200; ncu: 66 % XU, mio_throttle the top stall with 2 MUFU per sigmoid), so
// the RB denominators d_b = 1 + 2^(-z_b log2 e) share a single MUFU.RCP of their product and each
// 1/d_b is recovered with FMULs: 1.25 MUFU per sigmoid for RB = 4 instead of 2.  The exponent is
// clamped so that the product cannot overflow (sigmoid(z < -20.8) is returned as 2^-30 ~ 9e-10:
// absolute error < 1e-9, below fp32 resolution of the activations).
template <int RB>
__device__ __forceinline__ void sigmoid_group(const float (&z)[RB], float (&s)[RB]) {
    constexpr float kClamp = RB >= 4 ? 30.0f : 60.0f;
    float d[RB];
#pragma unroll
    for (int b = 0; b < RB; ++b) d[b] = 1.0f + ex2_ftz(fminf(-1.4426950408889634f * z[b], kClamp));
    if constexpr (RB == 4) {
        const float p01 = d[0] * d[1], p23 = d[2] * d[3];
        const float r = rcp_ftz(p01 * p23);
        const float r01 = r * p23, r23 = r * p01;
        s[0] = r01 * d[1]; s[1] = r01 * d[0]; s[2] = r23 * d[3]; s[3] = r23 * d[2];
    } else if constexpr (RB == 2) {
        const float r = rcp_ftz(d[0] * d[1]);
        s[0] = r * d[1]; s[1] = r * d[0];
    } else {
#pragma unroll
        for (int b = 0; b < RB; ++b) s[b] = rcp_ftz(d[b]);
    }
}

template <bool PRECISE>
__device__ __forceinline__ float sigmoid_sel(float z) {
    if constexpr (PRECISE) return sigmoid_precise(z);
    else return sigmoid_fast(z);
}

__device__ __forceinline__ float warp_sum_shfl(float v, int levels) {
    // xor butterfly over the lowest 2^levels lanes (callers with H <= 16 need fewer levels)
    if (levels >= 5) v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (levels >= 4) v += __shfl_xor_sync(0xffffffffu, v, 8);
    if (levels >= 3) v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (levels >= 2) v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (levels >= 1) v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// All-lane sum on the critical path of the SGD recurrence.  Measured on B200 (tools/latency_probe.cu):
// one SHFL.BFLY+FADD level costs ~36 cycles, so a 5-level butterfly is ~180 cycles per row, while
// FMUL + F2I + REDUX.SUM.S32 + I2F + FMUL is ~75.  The sum is therefore taken in 2^-22 fixed point
// with the integer REDUX unit (exact and order-independent once quantised; quantisation error
// <= 32 * 2^-23 ~ 4e-6 worst case, ~4e-7 rms, i.e. at the level of fp32 summation error).  A second,
// concurrent REDUX.MAX on the magnitudes guards the range: if any lane holds |v| >= 15.9 the
// butterfly is used instead (warp-uniform branch; needs |weight| ~ 16, far outside the prior's bulk).
__device__ __forceinline__ float warp_sum(float v, int levels) {
    if (levels <= 1) return warp_sum_shfl(v, levels);
    const unsigned int mag = __reduce_max_sync(0xffffffffu, __float_as_uint(v) & 0x7fffffffu);
    const int q = __float2int_rn(v * 4194304.0f);
    const int s = __reduce_add_sync(0xffffffffu, q);
    if (mag < 0x417e6666u)   // 15.9f
        return (float)s * (1.0f / 4194304.0f);
    return warp_sum_shfl(v, 5);
}

// O sums at once: the O integer REDUX operations are independent (they pipeline), and ONE range
// guard covers all of them, so there is a single warp-uniform branch per row instead of O.
template <int O>
__device__ __forceinline__ void warp_sum_vec(float (&v)[O], int levels) {
    if (levels <= 1) {
#pragma unroll
        for (int o = 0; o < O; ++o) v[o] = warp_sum_shfl(v[o], levels);
        return;
    }
    unsigned int m = 0u;
#pragma unroll
    for (int o = 0; o < O; ++o) m = max(m, __float_as_uint(v[o]) & 0x7fffffffu);
    const unsigned int mag = __reduce_max_sync(0xffffffffu, m);
    int s[O];
#pragma unroll
    for (int o = 0; o < O; ++o) s[o] = __reduce_add_sync(0xffffffffu, __float2int_rn(v[o] * 4194304.0f));
    if (mag < 0x417e6666u) {   // 15.9f
#pragma unroll
        for (int o = 0; o < O; ++o) v[o] = (float)s[o] * (1.0f / 4194304.0f);
    } else {
#pragma unroll
        for (int o = 0; o < O; ++o) v[o] = warp_sum_shfl(v[o], 5);
    }
}

__device__ __forceinline__ unsigned int hw_smid() { unsigned int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned int hw_warpid() { unsigned int r; asm volatile("mov.u32 %0, %%warpid;" : "=r"(r)); return r; }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of K doubles.  `scratch` holds K * (NT/32) doubles.  Result valid in ALL threads.
template <int K, int NT>
__device__ __forceinline__ void block_sum(double (&v)[K], double *scratch) {
    constexpr int NW = NT / kWarp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum_d(v[k]);
    __syncthreads();  // scratch may still be read from a previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) scratch[k * NW + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += scratch[k * NW + w];  // fixed order: deterministic
        v[k] = s;
    }
}

// ------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier: stage the training/test sets into shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// The same wait for a warp that may wait long next to warps that are busy (the MMA warp of the tcgen05 pass): the
// thread is SUSPENDED by the hardware until the phase completes or the hint (ns) runs out, instead of re-issuing
// try_wait every few cycles from the issue slots of the warps that share its scheduler.
__device__ __forceinline__ void mbar_wait_suspended(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
    } while (!ok);
}
// 1-D TMA: global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// grid-wide barrier for the swap round (cooperative launch guarantees co-residency).
// Replaces the reference's per-replica Event pair (R:432-434, R:730-752).
// ------------------------------------------------------------------------------------------
struct GridBarrier {
    unsigned int count;
    unsigned int generation;
    unsigned int failed;      // set when a CTA gave up waiting (the grid was not co-resident): reported by the host
    unsigned int pad;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// A wait that cannot hang the device.  The limit counts time WITHOUT PROGRESS, not waiting time: every CTA
// bumps a heartbeat word after each MCMC step, and a waiter restarts its clock whenever the heartbeat it
// watches has moved (legitimate skew -- a CTA that owns one more replica segment than its neighbours, a peer
// rank that started late -- is progress, a grid that is not co-resident is not).  `limit` comes from the
// configuration (ptfnn_config::barrier_timeout_ms).
struct WaitClock {
    long long t0, limit;
    const unsigned int *beat;
    unsigned int last;
    __device__ __forceinline__ WaitClock(const unsigned int *heartbeat, long long limit_cycles) {
        t0 = clock64(); limit = limit_cycles; beat = heartbeat;
        last = heartbeat ? ld_relaxed_u32(heartbeat) : 0u;
    }
    // true: give up
    __device__ __forceinline__ bool expired() {
        const long long now = clock64();
        if (now - t0 <= limit) return false;
        if (beat) {
            const unsigned int b = ld_relaxed_u32(beat);
            if (b != last) { last = b; t0 = now; return false; }
        }
        return true;
    }
};

// Returns false -- in EVERY thread of the CTA -- when this barrier (or an earlier one) timed out: the caller
// must leave the kernel without touching chain state, traces or the swap window.
__device__ __forceinline__ bool grid_barrier(GridBarrier *b, unsigned int nblocks, long long limit_cycles,
                                             const unsigned int *heartbeat, bool spin = false) {
    int bad = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int gen = ld_acquire_u32(&b->generation);
        __threadfence();
        if (atomicAdd(&b->count, 1u) == nblocks - 1u) {
            b->count = 0u;
            __threadfence();
            atomicAdd(&b->generation, 1u);
        } else {
            WaitClock wc(heartbeat, limit_cycles);
            while (ld_acquire_u32(&b->generation) == gen) {
                if (!spin) __nanosleep(32);          // small groups (speculative windows) poll without backing off
                if (ld_relaxed_u32(&b->failed) || wc.expired()) { atomicExch(&b->failed, 1u); break; }   // never hang the device
            }
        }
        __threadfence();
        bad = (int)ld_relaxed_u32(&b->failed);
    }
    return __syncthreads_or(bad) == 0;
}

// ------------------------------------------------------------------------------------------
// peer memory (NVLink / NVSwitch): system-scope stores, loads and flags for the multi-GPU swap round
// ------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;
__device__ __forceinline__ void st_release_sys_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f32(const float *p) {      // never served from a stale L1 line
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_sys_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// eta crosses the swap window as the bit pattern of the float64 (the reference moves the float64, R:430-437)
__device__ __forceinline__ void put_f64_words(float *two_words, double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    unsigned int *t = reinterpret_cast<unsigned int *>(two_words);
    t[0] = (unsigned int)b; t[1] = (unsigned int)(b >> 32);
}
__device__ __forceinline__ double f64_from_words(unsigned int lo, unsigned int hi) {
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
__device__ __forceinline__ double ld_sys_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------
// swap rule (R:674): min(1, 0.5 * exp(min(709, l2 - l1))); sequential sweep (R:741-748).
// One thread; lh/src live in shared memory.  Returns the number of accepted swaps.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double swap_probability(double l1, double l2) {
    double d = l2 - l1;
    if (d > 709.0) d = 709.0;
    const double p = 0.5 * exp(d);
    return p < 1.0 ? p : 1.0;  // NaN -> 1, like Python's min(1, nan)
}

// Opt-in swap rule of the reference's drafts (Misc/ldpt_fnn_multi_fixed.py:520):
//     swap_proposal = (lhood1 / (1 if lhood2 == 0 else lhood2)) * (1/T1 * 1/T2),   swap when u < swap_proposal
// with T1, T2 the temperature fields that travel with the two vectors.  Not used by any published result
// of the reference and excluded from replay parity; kSwapKindReference is R:674.
constexpr int kSwapKindReference = 0;
constexpr int kSwapKindRatioTemp = 1;
__device__ __forceinline__ double swap_probability_ratio(double l1, double l2, double t1, double t2) {
    return (l1 / (l2 == 0.0 ? 1.0 : l2)) * (1.0 / t1 * 1.0 / t2);
}

template <class UFn, class TFn>
__device__ __forceinline__ int swap_sweep_serial(int n, double *lh, int *src, uint8_t *swapped_out, UFn u_of, int kind, TFn t_of) {
    int ns = 0;
    for (int k = 0; k + 1 < n; ++k) {
        const double pr = kind == kSwapKindReference ? swap_probability(lh[k], lh[k + 1])
                                                     : swap_probability_ratio(lh[k], lh[k + 1], t_of(src[k]), t_of(src[k + 1]));
        const bool s = (double)u_of(k) < pr;
        if (s) {
            const double t = lh[k]; lh[k] = lh[k + 1]; lh[k + 1] = t;
            const int ti = src[k]; src[k] = src[k + 1]; src[k + 1] = ti;
            ++ns;
        }
        if (swapped_out) swapped_out[k] = (uint8_t)s;
    }
    return ns;
}

// Streaming form of the same sweep with O(1) state: when pair k is decided, slot k's final content
// is known (either the original vector of slot k+1, or the vector that has been bubbling up), so
// nothing but the bubbling vector's (lhood, origin) needs to be carried.  `lh_chunk(k)` returns the
// ORIGINAL lhood of slot k; emit(slot, origin) reports final contents, decided(k, swapped) every pair.
// The decision u < min(1, 0.5 exp(min(709, d))) is taken in the log domain, log(2u) < min(709, d), with
// log(2u) precomputed by the whole CTA (`lu_of`): the sweep itself is sequential (pair k needs the outcome
// of pair k-1) and an fp64 exp per pair made it ~0.2 us per rung -- 1.7 ms per round on an 8192-rung ladder.
// Near a tie (and for u = 0 or a NaN difference) the reference's own expression decides, so the result is
// the reference's in every case.
__device__ __forceinline__ bool swap_decision(double cur_l, double nxt, float u, double lu) {
    double d = nxt - cur_l;
    if (d > 709.0) d = 709.0;
    const double tol = 1e-12 * fmax(1.0, fabs(d));
    if (lu < d - tol && lu > -1e300) return true;
    if (lu > d + tol) return false;
    return (double)u < swap_probability(cur_l, nxt);
}

template <class LFn, class UFn, class LUFn, class Emit, class Decided, class TFn>
__device__ __forceinline__ void swap_sweep_stream(int k_begin, int k_end, int n, double &cur_l, int &cur_src, int &ns,
                                                  LFn lh_of, UFn u_of, LUFn lu_of, Emit emit, Decided decided,
                                                  int kind, TFn t_of) {
    for (int k = k_begin; k < k_end && k + 1 < n; ++k) {
        const double nxt = lh_of(k + 1);
        const bool s = kind == kSwapKindReference ? swap_decision(cur_l, nxt, u_of(k), lu_of(k))
                                                  : (double)u_of(k) < swap_probability_ratio(cur_l, nxt, t_of(cur_src), t_of(k + 1));
        if (s) { emit(k, k + 1); ++ns; }
        else { emit(k, cur_src); cur_l = nxt; cur_src = k + 1; }
        decided(k, s);
    }
}
#endif  // __CUDACC__

}  // namespace ptfnn
