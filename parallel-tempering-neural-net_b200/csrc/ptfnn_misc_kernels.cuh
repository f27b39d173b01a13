// Topology-independent kernels (swap sweep, prior, Philox dump, multi-GPU row install).
// Included by ptfnn_api.cu only (non-template __global__ definitions).
#pragma once
#include "ptfnn_device.cuh"

namespace ptfnn {

// ---- topology-independent kernels ----
__global__ void op_sweep_kernel(int n, const double *lhood, const float *u, int *src, uint8_t *swapped, int *ns_out,
                                int kind, const double *temperature) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *l = reinterpret_cast<double *>(smem_raw);
    int *s = reinterpret_cast<int *>(smem_raw + (size_t)n * 8);
    for (int k = threadIdx.x; k < n; k += blockDim.x) { l[k] = lhood[k]; s[k] = k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ns = swap_sweep_serial(n, l, s, swapped, [&](int k) { return u[k]; }, kind,
                                         [&](int slot) { return temperature ? temperature[slot] : 1.0; });
        if (ns_out) *ns_out = ns;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) src[k] = s[k];
}

__global__ void op_prior_kernel(int task, int I, int H, int O, const float *w, double sigma_sq, double nu1,
                                double nu2, double tausq, double *out) {
    __shared__ double red[32];
    const int P = I * H + H * O + H + O;
    double s = 0.0;
    for (int j = threadIdx.x; j < P; j += blockDim.x) s += (double)w[j] * (double)w[j];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
        if (task == kTaskReg)
            *out = -((I * H + H + 2) / 2.0) * log(sigma_sq) - t / (2.0 * sigma_sq) - (1.0 + nu1) * log(tausq) - nu2 / tausq;
        else
            *out = -((I * H + H + O + H * O) / 2.0) * log(sigma_sq) - t / (2.0 * sigma_sq);
    }
}

// Feedback sample for the automatic depth of the speculative windows: out[0] = sum of the replicas' acceptance
// counters, out[1] / out[2] = Langevin / random-walk steps among the n steps the launch just made (first replica's
// stream; lx != nullptr: replayed draws).
__global__ void feedback_kernel(const int *n_acc, int R, uint64_t seed, int step0, int n, uint32_t stream, uint32_t gr,
                                const float *lx, int use_lg, double l_prob, long long *out) {
    __shared__ long long red[8][2];
    long long a = 0, g = 0;
    for (int k = threadIdx.x; k < R; k += blockDim.x) a += n_acc[k];
    if (use_lg)
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const float x = lx ? lx[k] : philox_step_scalars(seed, (uint32_t)(step0 + k), stream, gr).lx;
            g += ((double)x < l_prob) ? 1 : 0;
        }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); g += __shfl_xor_sync(0xffffffffu, g, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = a; red[threadIdx.x >> 5][1] = g; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0, u = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { t += red[k][0]; u += red[k][1]; }
        out[0] = t; out[1] = u; out[2] = n - u;
    }
}

// Dump the Philox draws of free-running mode (verification only).
__global__ void draws_kernel(uint64_t seed, int crn, int replica_offset, int R, int P, int i0, int n, float *lx,
                             float *z, float *z_eta, float *u) {
    const int r = blockIdx.y, k = blockIdx.x;   // replica, step offset
    const uint32_t gr = (uint32_t)(replica_offset + r);
    const uint32_t stream = crn ? kStreamCommon : gr;
    const size_t q = (size_t)r * n + k;
    if (threadIdx.x == 0) {
        const StepDraws sd = philox_step_scalars(seed, (uint32_t)(i0 + k), stream, gr);
        lx[q] = sd.lx; z_eta[q] = sd.z_eta; u[q] = sd.u;
    }
    for (int b = threadIdx.x; b < (P + 3) / 4; b += blockDim.x) {
        float z4[4];
        philox_step_normals4(seed, (uint32_t)(i0 + k), stream, (uint32_t)b, z4);
        for (int c = 0; c < 4; ++c)
            if (4 * b + c < P) z[q * P + 4 * b + c] = z4[c];
    }
}

// Network.ForwardPass on ONE row with runtime dimensions (R:51-55): hid = sigmoid(x.W1 - B1),
// out = sigmoid(hid.W2 - B2).  One block; hidden units strided over threads.
__global__ void op_forward_row_kernel(int I, int H, int O, const float *x, const float *w, float *hid, float *out) {
    extern __shared__ float s_hid[];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float z = -w[I * H + H * O + h];
        for (int i = 0; i < I; ++i) z = fmaf(x[i], w[i * H + h], z);
        s_hid[h] = sigmoid_precise(z);
        hid[h] = s_hid[h];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        float z = -w[I * H + H * O + H + o];
        for (int h = 0; h < H; ++h) z = fmaf(s_hid[h], w[I * H + h * O + o], z);
        out[o] = sigmoid_precise(z);
    }
}

// ---- result pipeline on the device traces ----
// The reference slices every trace at the burn-in row and pools chains x samples (R:775-871); its main()
// then reports mean / np.std / min of the pooled RMSE columns (R:1036-1044).  trace_summary_kernel does
// that reduction where the traces already are, in ONE launch, so a summary costs one read of the trace at
// HBM speed instead of the device->host copy (and the txt round trip) of R x S x P values.
//
// Blocks of the (1-D) grid: ctiles planes of mblocks persistent blocks (or one or two short blocks per SM when only
// the scalar series are wanted).  Every block does two things:
//   the scalar series  rmse_train, rmse_test, acc_train, acc_test (fp64 [R, S]): the block strides over the
//                      replicas, reads the four series side by side (four independent loads per thread and
//                      step) and leaves its partial {sum, sum of squares, min, max} of each.  This is 32 bytes per
//                      pooled row against 4P for the weights, folded into the streaming blocks so that it
//                      costs neither extra waves of blocks behind them nor block slots in front of them;
//   the posterior moments (plane = column tile)  -- for every parameter p, sum and sum of
//                      squares of (pos_w[r, first+i, p] - pivot[p]) over all local replicas r and rows i < count;
//                      pivot[p] = pos_w[0, first, p] keeps the second moment free of cancellation (np.std is
//                      two-pass).  HBM-bound: R*count*P floats read exactly once; fp64 sums.
//                      Persistent blocks (2 per SM) stream the trace through a 3-stage ring of 32 KB shared-
//                      memory buffers filled by 1-D TMA bulk copies (~190 KB in flight per SM, no registers
//                      tied up by loads); the 256 threads then read a stage conflict-free with FIXED columns
//                      per thread, so there is no index arithmetic per element.  The burn-in slice of one
//                      replica is a contiguous span of count*P floats: when one column tile covers the row
//                      (P <= 2048) a stage is one copy of as many whole rows as fit, and P < 256 packs
//                      `groups` = 256/P adjacent rows into each pass of the block; wider nets tile the
//                      columns by 2048 (one plane per tile) and a stage holds four row pieces.  Rows are not
//                      16-byte aligned (P is odd as a rule): every copy starts at the aligned address below
//                      its first element and the reader skips the `shift` floats in front.
// The last block to finish (ticket counter) folds everything into {mean, np.std, min, max} per series and
// mean / np.std per parameter, then clears the accumulators for the next call.
constexpr int kSumThreads = 256;
constexpr int kSumStages = 3;
constexpr int kSumMaxCols = 8;
constexpr int kSumPad = 8;                                                       // slack per staged piece (alignment at both ends)
constexpr int kSumStageFloats = 4 * (kSumThreads * kSumMaxCols + kSumPad);       // 8224 floats
constexpr size_t kSumSmemBytes = (size_t)kSumStages * kSumStageFloats * sizeof(float) + kSumStages * sizeof(uint64_t);
constexpr int kSumTracePadFloats = 4;    // the trace allocation ends with this much padding: aligned copies may run past the last row

struct TraceSummaryArgs {
    const float *pos_w;        // [R, S, P] (+ kSumTracePadFloats)
    const double *series[4];   // [R, S] each
    int R, S, P, first, count, ctiles;
    int mblocks;               // blocks per column tile: gridDim.x = ctiles*mblocks (any number when ctiles == 0)
    int with_series;           // 0: moments only (the matrix is not a trace: ptfnn_predictive_summary)
    double *acc;               // [2][P], zero on entry, zero on exit
    double *part;              // [gridDim.x][4 series][4]
    unsigned int *ticket;      // zero on entry, zero on exit
    double *stats;             // out [4][4]
    double *mean, *stdev;      // out [P] (ctiles > 0)
};

__device__ __forceinline__ const double *series_of(const TraceSummaryArgs &a, int k) {   // (no dynamic indexing of a kernel parameter)
    return k == 0 ? a.series[0] : k == 1 ? a.series[1] : k == 2 ? a.series[2] : a.series[3];
}

__device__ __forceinline__ void fold4(double &s1, double &s2, double &lo, double &hi) {
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
}

// rows of one stage: whole rows back to back, or row pieces at a fixed pitch
__host__ __device__ inline int sum_rows_per_chunk(int P, int cols, bool whole) {
    const int n = whole ? (kSumStageFloats - kSumPad) / P : kSumStageFloats / (kSumThreads * cols + kSumPad);
    return n < 1 ? 1 : n;
}

template <int COLS>
__global__ void __launch_bounds__(kSumThreads, 2) trace_summary_kernel(const TraceSummaryArgs a) {
    extern __shared__ __align__(128) unsigned char sum_smem[];
    __shared__ double s_fold[16][kSumThreads / 32];
    __shared__ bool s_last;
    const int P = a.P, count = a.count, tid = threadIdx.x;
    if (a.ctiles > 0) {
        const int plane = (int)blockIdx.x / a.mblocks, bx = (int)blockIdx.x - plane * a.mblocks, nbx = a.mblocks;
        constexpr int CW = kSumThreads * COLS, PITCH = CW + kSumPad;
        float *stage = reinterpret_cast<float *>(sum_smem);
        uint64_t *full = reinterpret_cast<uint64_t *>(sum_smem + (size_t)kSumStages * kSumStageFloats * sizeof(float));
        const bool whole = a.ctiles == 1;
        const int c0 = plane * CW, cw = min(P - c0, CW);
        const int width = min(cw, kSumThreads);
        const int groups = whole ? kSumThreads / width : 1;
        const int g = tid / width, c = tid - g * width;
        const int rpc = sum_rows_per_chunk(P, COLS, whole);
        const int chunks_per_rep = (count + rpc - 1) / rpc;
        const int items = a.R * chunks_per_rep;
        const int n_mine = bx < items ? (items - bx + nbx - 1) / nbx : 0;
        double s1[COLS], s2[COLS], pivot[COLS];
        bool live[COLS];
#pragma unroll
        for (int k = 0; k < COLS; ++k) {
            live[k] = g < groups && c + k * kSumThreads < cw;
            s1[k] = s2[k] = 0.0;
            pivot[k] = live[k] ? (double)__ldg(a.pos_w + (size_t)a.first * P + c0 + c + k * kSumThreads) : 0.0;
        }
        auto locate = [&](int n, int &r, int &i0, int &nrows) {
            const int t = bx + n * nbx;
            r = t / chunks_per_rep;
            i0 = (t - r * chunks_per_rep) * rpc;
            nrows = min(rpc, count - i0);
        };
        auto issue = [&](int n) {                                   // one thread: fill stage n % kSumStages with item n
            int r, i0, nrows;
            locate(n, r, i0, nrows);
            const int slot = n % kSumStages;
            float *dst = stage + (size_t)slot * kSumStageFloats;
            const size_t row0 = (size_t)r * a.S + a.first + i0;
            if (whole) {
                const size_t e0 = row0 * P, e1 = e0 + (size_t)nrows * P;
                const size_t b0 = e0 & ~(size_t)3, b1 = (e1 + 3) & ~(size_t)3;
                const uint32_t bytes = (uint32_t)(b1 - b0) * 4u;
                mbar_arrive_expect_tx(&full[slot], bytes);
                tma_load_1d(dst, a.pos_w + b0, bytes, &full[slot]);
            } else {
                uint32_t bytes = 0;
                for (int i = 0; i < nrows; ++i) {
                    const size_t e0 = (row0 + i) * P + c0, e1 = e0 + cw;
                    bytes += (uint32_t)(((e1 + 3) & ~(size_t)3) - (e0 & ~(size_t)3)) * 4u;
                }
                mbar_arrive_expect_tx(&full[slot], bytes);
                for (int i = 0; i < nrows; ++i) {
                    const size_t e0 = (row0 + i) * P + c0, e1 = e0 + cw;
                    const size_t b0 = e0 & ~(size_t)3, b1 = (e1 + 3) & ~(size_t)3;
                    tma_load_1d(dst + (size_t)i * PITCH, a.pos_w + b0, (uint32_t)(b1 - b0) * 4u, &full[slot]);
                }
            }
        };
        if (tid == 0) {
            for (int q = 0; q < kSumStages; ++q) mbar_init(&full[q], 1);
            mbar_fence_init();
            for (int n = 0; n < min(kSumStages, n_mine); ++n) issue(n);
        }
        __syncthreads();
        for (int n = 0; n < n_mine; ++n) {
            int r, i0, nrows;
            locate(n, r, i0, nrows);
            const int slot = n % kSumStages;
            const float *src = stage + (size_t)slot * kSumStageFloats + c;
            const size_t row0 = (size_t)r * a.S + a.first + i0;
            mbar_wait(&full[slot], (uint32_t)(n / kSumStages) & 1u);
            if (whole) {
                const float *p = src + (int)((row0 * P) & 3) + g * P;           // row j*groups + g, columns c + 256k
                const int stride = groups * P;
                const int full_passes = nrows / groups;
                int j = 0;
                for (; j + 4 <= full_passes; j += 4) {
                    float v[4][COLS];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int k = 0; k < COLS; ++k) v[u][k] = live[k] ? p[(j + u) * stride + k * kSumThreads] : 0.0f;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int k = 0; k < COLS; ++k) {
                            const double d = (double)v[u][k] - pivot[k];
                            s1[k] += d; s2[k] = fma(d, d, s2[k]);
                        }
                }
                for (; j * groups + g < nrows; ++j)
#pragma unroll
                    for (int k = 0; k < COLS; ++k)
                        if (live[k]) {
                            const double d = (double)p[j * stride + k * kSumThreads] - pivot[k];
                            s1[k] += d; s2[k] = fma(d, d, s2[k]);
                        }
            } else {
                for (int i = 0; i < nrows; ++i) {
                    const float *p = src + (size_t)i * PITCH + (int)(((row0 + i) * P + c0) & 3);
#pragma unroll
                    for (int k = 0; k < COLS; ++k)
                        if (live[k]) {
                            const double d = (double)p[k * kSumThreads] - pivot[k];
                            s1[k] += d; s2[k] = fma(d, d, s2[k]);
                        }
                }
            }
            __syncthreads();                                            // the stage is drained: refill it
            if (tid == 0 && n + kSumStages < n_mine) issue(n + kSumStages);
        }
        if (groups == 1) {
#pragma unroll
            for (int k = 0; k < COLS; ++k)
                if (live[k]) {
                    atomicAdd(a.acc + c0 + c + k * kSumThreads, s1[k]);
                    atomicAdd(a.acc + P + c0 + c + k * kSumThreads, s2[k]);
                }
        } else {                                                        // P < 256: fold the row groups first (COLS = 1)
            double *red = reinterpret_cast<double *>(sum_smem);         // every copy has landed and been read
            red[tid] = s1[0]; red[kSumThreads + tid] = s2[0];
            __syncthreads();
            if (g == 0) {
                double x = 0.0, y = 0.0;
                for (int gg = 0; gg < groups; ++gg) { x += red[gg * width + c]; y += red[kSumThreads + gg * width + c]; }
                atomicAdd(a.acc + c, x);
                atomicAdd(a.acc + P + c, y);
            }
        }
    }
    if (a.with_series) {   // ---- the four scalar series, side by side
        const double *x0 = a.series[0], *x1 = a.series[1], *x2 = a.series[2], *x3 = a.series[3];
        double pv[4] = {x0[a.first], x1[a.first], x2[a.first], x3[a.first]};
        double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
        double lo[4] = {pv[0], pv[1], pv[2], pv[3]}, hi[4] = {pv[0], pv[1], pv[2], pv[3]};
        for (int r = blockIdx.x; r < a.R; r += gridDim.x) {
            const size_t off = (size_t)r * a.S + a.first;
            for (int i = tid; i < count; i += kSumThreads) {
                const double v[4] = {__ldcs(x0 + off + i), __ldcs(x1 + off + i), __ldcs(x2 + off + i), __ldcs(x3 + off + i)};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double d = v[q] - pv[q];
                    s1[q] += d; s2[q] = fma(d, d, s2[q]);
                    lo[q] = fmin(lo[q], v[q]); hi[q] = fmax(hi[q], v[q]);
                }
            }
        }
        const int warp = tid >> 5;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            fold4(s1[q], s2[q], lo[q], hi[q]);
            if ((tid & 31) == 0) { s_fold[4 * q][warp] = s1[q]; s_fold[4 * q + 1][warp] = s2[q]; s_fold[4 * q + 2][warp] = lo[q]; s_fold[4 * q + 3][warp] = hi[q]; }
        }
        __syncthreads();
        if (tid < 4) {
            const int q = tid;
            double t1 = 0.0, t2 = 0.0, tl = s_fold[4 * q + 2][0], th = s_fold[4 * q + 3][0];
            for (int k = 0; k < kSumThreads / 32; ++k) {
                t1 += s_fold[4 * q][k]; t2 += s_fold[4 * q + 1][k]; tl = fmin(tl, s_fold[4 * q + 2][k]); th = fmax(th, s_fold[4 * q + 3][k]);
            }
            double *o = a.part + ((size_t)blockIdx.x * 4 + q) * 4;
            o[0] = t1; o[1] = t2; o[2] = tl; o[3] = th;
        }
    }
    // ---- the last block folds the partial results
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double n = (double)a.R * (double)count;
    if (tid < 128 && a.with_series) {
        const int sidx = tid >> 5, lane = tid & 31;
        const double pivot = series_of(a, sidx)[a.first];
        double s1 = 0.0, s2 = 0.0, lo = pivot, hi = pivot;
        for (int b = lane; b < (int)gridDim.x; b += 32) {
            const double *o = a.part + ((size_t)b * 4 + sidx) * 4;
            s1 += __ldcg(o); s2 += __ldcg(o + 1); lo = fmin(lo, __ldcg(o + 2)); hi = fmax(hi, __ldcg(o + 3));
        }
        fold4(s1, s2, lo, hi);
        if (lane == 0) {
            const double m = s1 / n, var = s2 / n - m * m;
            double *o = a.stats + 4 * sidx;
            o[0] = pivot + m; o[1] = sqrt(var > 0.0 ? var : 0.0); o[2] = lo; o[3] = hi;
        }
    }
    if (a.ctiles > 0)
        for (int p = tid; p < P; p += kSumThreads) {
            const double m = __ldcg(a.acc + p) / n, var = __ldcg(a.acc + P + p) / n - m * m;
            a.mean[p] = (double)a.pos_w[(size_t)a.first * P + p] + m;
            a.stdev[p] = sqrt(var > 0.0 ? var : 0.0);
            a.acc[p] = 0.0; a.acc[P + p] = 0.0;
        }
    if (tid == 0) *a.ticket = 0u;
}

// ---- percentile bands of the posterior-predictive distribution ----
// m is the [n, N] prediction matrix (n pooled posterior samples, N data rows, float32).  For every data row the
// order statistics k0, k0+1, k1, k1+1 of its column are found EXACTLY: a most-significant-digit radix select
// (four 8-bit digits of the order-preserving integer image of the float, one histogram pass per digit, two
// targets at once), then one more pass that resolves the successor of each selected value (the next larger
// element, or the same value when it is tied across the two ranks).  A CTA owns kQCols adjacent columns, so a
// warp reads whole 32-byte sectors; the matrix is read five times and is L2-resident for the reference's
// configurations (25 000 samples x 298 rows = 30 MB).  lo / hi follow np.percentile's default (linear) rule.
constexpr int kQCols = 8;
constexpr int kQThreads = 256;
__device__ __forceinline__ unsigned int f32_ordered(float v) {
    const unsigned int b = __float_as_uint(v);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float f32_from_ordered(unsigned int k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}
__global__ void __launch_bounds__(kQThreads) quantile_bands_kernel(const float *__restrict__ m, long long n, int N, long long k0, double f0,
                                                                   long long k1, double f1, double *lo, double *hi) {
    __shared__ unsigned int hist[2][256][kQCols];
    __shared__ unsigned int prefix[2][kQCols];
    __shared__ unsigned long long rank[2][kQCols];
    __shared__ unsigned long long cnt_le[2][kQCols];
    __shared__ unsigned int min_gt[2][kQCols];
    const int tid = threadIdx.x, c = tid % kQCols, sl = tid / kQCols;
    constexpr int SL = kQThreads / kQCols;
    const int col = blockIdx.x * kQCols + c;
    const bool live = col < N;
    if (tid < 2 * kQCols) { prefix[tid / kQCols][tid % kQCols] = 0u; rank[tid / kQCols][tid % kQCols] = (unsigned long long)(tid / kQCols == 0 ? k0 : k1); }
    for (int d = 3; d >= 0; --d) {
        for (int i = tid; i < 2 * 256 * kQCols; i += kQThreads) (&hist[0][0][0])[i] = 0u;
        __syncthreads();
        const unsigned int p0 = prefix[0][c], p1 = prefix[1][c];
        const int sh = 8 * d;
        const bool same = p0 == p1;                      // both targets still inside the same bucket: one histogram serves both
        if (live)
            for (long long r = sl; r < n; r += SL) {
                const unsigned int key = f32_ordered(m[r * N + col]);
                const unsigned int hi_bits = d == 3 ? 0u : (key >> (sh + 8));
                const unsigned int dg = (key >> sh) & 255u;
                if (d == 3 || hi_bits == (p0 >> (sh + 8))) atomicAdd(&hist[0][dg][c], 1u);
                if (!same && hi_bits == (p1 >> (sh + 8))) atomicAdd(&hist[1][dg][c], 1u);
            }
        __syncthreads();
        if (tid < 2 * kQCols) {
            const int t = tid / kQCols, cc = tid % kQCols;
            const int h = (prefix[0][cc] == prefix[1][cc]) ? 0 : t;
            unsigned long long want = rank[t][cc], cum = 0;
            int dsel = 255;
            for (int b = 0; b < 256; ++b) {
                const unsigned int v = hist[h][b][cc];
                if (cum + v > want) { dsel = b; break; }
                cum += v;
            }
            rank[t][cc] = want - cum;
            __syncwarp((1u << (2 * kQCols)) - 1u);       // both prefixes of a column were read (h) before either is updated
            prefix[t][cc] |= (unsigned int)dsel << sh;
        }
        __syncthreads();
    }
    // successor of each selected value
    if (tid < 2 * kQCols) { cnt_le[tid / kQCols][tid % kQCols] = 0ull; min_gt[tid / kQCols][tid % kQCols] = 0xffffffffu; }
    __syncthreads();
    {
        const unsigned int v0 = prefix[0][c], v1 = prefix[1][c];
        unsigned long long le0 = 0, le1 = 0;
        unsigned int g0 = 0xffffffffu, g1 = 0xffffffffu;
        if (live)
            for (long long r = sl; r < n; r += SL) {
                const unsigned int key = f32_ordered(m[r * N + col]);
                le0 += key <= v0; le1 += key <= v1;
                if (key > v0) g0 = min(g0, key);
                if (key > v1) g1 = min(g1, key);
            }
        atomicAdd(&cnt_le[0][c], le0); atomicAdd(&cnt_le[1][c], le1);
        atomicMin(&min_gt[0][c], g0); atomicMin(&min_gt[1][c], g1);
    }
    __syncthreads();
    if (tid < kQCols && blockIdx.x * kQCols + tid < N) {
        const int cc = tid;
        auto band = [&](int t, long long k, double f) {
            const double a = (double)f32_from_ordered(prefix[t][cc]);
            // rank k+1 holds the same value when at least k+2 elements are <= it, else the next larger element
            const bool tied = cnt_le[t][cc] >= (unsigned long long)(k + 2) || min_gt[t][cc] == 0xffffffffu;
            const double b = tied ? a : (double)f32_from_ordered(min_gt[t][cc]);
            return f > 0.0 ? a + (b - a) * f : a;
        };
        lo[blockIdx.x * kQCols + cc] = band(0, k0, f0);
        hi[blockIdx.x * kQCols + cc] = band(1, k1, f1);
    }
}

// multi-GPU: install the rows selected by the sweep (R:435-437)
__global__ void swap_apply_kernel(int R, int P, int replica_offset, const int *src, const float *rows_local,
                                  const float *rows_in, float *w, double *eta, int *gd_valid) {
    const int r = blockIdx.x;
    const int s = src[replica_offset + r];
    if (s == replica_offset + r) return;
    const bool local = s >= replica_offset && s < replica_offset + R;
    constexpr int kTail = 2;                    // eta travels as its fp64 bit pattern (ptfnn_kernels.cuh: kRowTail)
    const float *row = local ? rows_local + (size_t)(s - replica_offset) * (P + kTail) : rows_in + (size_t)r * (P + kTail);
    for (int j = threadIdx.x; j < P; j += blockDim.x) w[(size_t)r * P + j] = row[j];
    if (threadIdx.x == 0) {
        const unsigned int *t = reinterpret_cast<const unsigned int *>(row + P);
        eta[r] = __longlong_as_double((long long)(((unsigned long long)t[1] << 32) | t[0]));
        gd_valid[r] = 0;
    }
}

}  // namespace ptfnn
