// libptfnn.so -- host side of the C ABI declared in include/ptfnn.h.
// Owns device memory, picks the kernel specialisation for the topology, launches the persistent
// chain kernel (cooperatively, one CTA per temperature) and copies traces back.  No CPU compute
// path exists: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/ptfnn.h"
#include "ptfnn_kernels.cuh"
#include "ptfnn_misc_kernels.cuh"

using namespace ptfnn;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

struct ptfnn_sampler;
static int fail(ptfnn_sampler *s, int code, const char *fmt, ...);

#define CU_TRY(S, expr)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (expr);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail((S), PTFNN_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------
// kernel dispatch table: one row per specialised topology
// ------------------------------------------------------------------------------------------
#include "ptfnn_registry.h"
#include "ptfnn_topologies.h"
typedef PtfnnKernelSet KernelSet;
static const int kMaxSpec = kMaxSpecK;          // most CTAs per temperature (speculative windows)

#define X(NAME, TASK, I, H, O, NT, MINB) const PtfnnKernelSet *ptfnn_kernelset_##NAME();
PTFNN_TOPOLOGIES(X)
#undef X

static std::vector<const KernelSet *> &kernel_sets() {
    static std::vector<const KernelSet *> v = {
#define X(NAME, TASK, I, H, O, NT, MINB) ptfnn_kernelset_##NAME(),
        PTFNN_TOPOLOGIES(X)
#undef X
    };
    return v;      // + specialisations registered at run time (ptfnn_register_kernels)
}

// handles may be created from several host threads while another one registers a specialisation
static std::mutex &kernel_sets_mutex() { static std::mutex m; return m; }

static const KernelSet *find_kernels(int task, int I, int H, int O) {
    std::lock_guard<std::mutex> g(kernel_sets_mutex());
    for (const KernelSet *k : kernel_sets())
        if (k->task == task && k->I == I && k->H == H && k->O == O) return k;
    return nullptr;
}

static std::string supported_list() {
    std::string s;
    char buf[64];
    std::lock_guard<std::mutex> g(kernel_sets_mutex());
    for (const KernelSet *k : kernel_sets()) {
        snprintf(buf, sizeof buf, "%s[%d,%d,%d] ", k->task == kTaskReg ? "reg" : "cls", k->I, k->H, k->O);
        s += buf;
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }          // error paths (CU_TRY returns) do not leak device memory
};

struct ptfnn_sampler {
    ptfnn_config cfg;
    const KernelSet *ks = nullptr;
    int P = 0, IP = 0;
    int swap_rule = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 0, regs_per_sm = 0, threads_per_sm = 0, clock_khz = 0;
    size_t smem_per_sm = 0;
    bool device_failed = false;       // a device-side wait gave up: sticky until ptfnn_init_chains
    unsigned int peer_round_base = 0; // flag domain of the peer hand-shake: grows with every ptfnn_init_chains
    int verified_grid = 0;            // TMEM kernels: CTAs proven co-resident by the probe launch ...
    size_t verified_smem = 0;         // ... for this much dynamic shared memory
    DevBuf<unsigned int> probe_count;
    std::string err;
    void *pinned = nullptr;           // host staging buffer of get_traces (page-locked, grows on demand)
    size_t pinned_bytes = 0;
    // overlapped read-back (ptfnn_traces_begin / _end): two page-locked slots filled on a copy stream
    struct FetchSlot { void *buf = nullptr; size_t bytes = 0; cudaEvent_t done = nullptr; size_t off[7] = {}; bool has_w = false; int first = 0, count = 0; bool busy = false; };
    FetchSlot slot[2];
    // ptfnn_set_data: page-locked staging of the packed data set, double buffered (no wait for the device)
    struct StageSlot { float *buf = nullptr; size_t bytes = 0; cudaEvent_t done = nullptr; bool used = false; };
    StageSlot stage[2];
    unsigned int stage_seq = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_compute = nullptr;
    unsigned int fetch_seq = 0;
    // acceptance feedback for the automatic depth of the speculative windows: after every launch the sum of the
    // replicas' acceptance counters is copied (asynchronously) into a small page-locked ring; the next launch looks at
    // the newest copy that has arrived, and waits for the copy of the launch before the previous one if it has not:
    // the host never runs more than two launches ahead of the estimate it steers by (a caller that queues a whole
    // run of launches at once would otherwise decide all of them blind)
    // The same sample times the launch (events around it) and counts its Langevin / random-walk steps: cost per
    // random-walk-equivalent step of the two "arms" -- sequential, windows -- is what finally picks between them.
    struct AccSample { long long *host = nullptr; cudaEvent_t ev = nullptr, t0 = nullptr, t1 = nullptr; int step = 0, arm = 0, launch = 0; bool used = false, timed = false; };
    AccSample acc_ring[4];
    DevBuf<long long> acc_sum;
    unsigned int acc_seq = 0;
    long long acc_prev_sum = 0; int acc_prev_step = 0; bool acc_have_prev = false;
    double acc_est = -1.0;            // < 0: nothing known yet
    double arm_cost[2] = {-1.0, -1.0};   // ms per random-walk-equivalent step: [0] sequential, [1] windows; < 0: unknown
    int arm_launched[2] = {-1, -1};      // feedback launch index at which the arm ran last
    double arm_acc[2] = {1.0, 1.0};      // acceptance estimate at that launch
    int arm_timed[2] = {-100, -100};     // launch index of the arm's last harvested timing
    int arm_interval[2] = {32, 32};      // launches between two tries of the arm while it loses (doubles every time it loses again)
    int fb_launches = 0;

    bool have_data = false, have_state = false, summary_smem_opted = false;
    int n_train = 0, n_test = 0;
    int step = 0, rounds_done = 0;
    bool swap_pending = false, pending_final = false;
    int64_t host_num_swap = 0, host_total_prop = 0;   // multi-GPU rounds are counted on the host
    std::vector<uint8_t> host_swap_log;               // [rounds][Rg-1] for externally planned rounds
    std::vector<int> host_swap_log_round;

    DevBuf<float> train_x, train_y, test_x, test_y;
    DevBuf<float> a_train, a_test;                    // K5: UMMA A tiles of the data sets (wide-hidden topologies)
    DevBuf<double> temperature;
    DevBuf<float> w, gd_cache, pgd_buf, prop_buf, pos_w, pub_rows;
    DevBuf<double> eta, tau, lik, prior, last4, init_rmse, pub_lhood;
    DevBuf<double> lik_prop, rmse_tr, rmse_te, acc_tr, acc_te, dbg_prior, dbg_diff, dbg_mh;
    DevBuf<int> n_acc, init_count, gd_valid, accept_list;
    DevBuf<uint8_t> dbg_acc, swap_log;
    DevBuf<GridBarrier> barrier, spec_bar;            // spec_bar: one per temperature (speculative windows)
    DevBuf<unsigned int> spec_flag;
    // multi-GPU ladder through peer memory (ptfnn_peer_connect)
    DevBuf<unsigned int> peer_flags;                  // [kMaxPeers] rounds published by each rank, [kMaxPeers] = this rank's heartbeat
    int n_ranks = 1, rank = 0;
    void *peer_lhood[kMaxPeers] = {}, *peer_rows[kMaxPeers] = {}, *peer_flag_ptr[kMaxPeers] = {};
    bool peer_opened[kMaxPeers][3] = {};
    DevBuf<long long> swap_counters;
    DevBuf<float> d_lx, d_z, d_zeta, d_u, d_uswap;   // replay staging
    DevBuf<int> d_src, smsp_load, swap_src;
    DevBuf<uint8_t> d_swapped;
    DevBuf<double> d_scratch, d_summary;

    void release_all() {
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr; pinned_bytes = 0;
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        for (auto &f : slot) {
            if (f.buf) cudaFreeHost(f.buf);
            if (f.done) cudaEventDestroy(f.done);
            f = FetchSlot();
        }
        for (auto &g : stage) {
            if (g.buf) cudaFreeHost(g.buf);
            if (g.done) cudaEventDestroy(g.done);
            g = StageSlot();
        }
        for (auto &a : acc_ring) {
            if (a.host) cudaFreeHost(a.host);
            if (a.ev) cudaEventDestroy(a.ev);
            if (a.t0) cudaEventDestroy(a.t0);
            if (a.t1) cudaEventDestroy(a.t1);
            a = AccSample();
        }
        acc_sum.release();
        if (ev_compute) cudaEventDestroy(ev_compute);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        ev_compute = nullptr; copy_stream = nullptr;
        train_x.release(); train_y.release(); test_x.release(); test_y.release(); a_train.release(); a_test.release(); temperature.release();
        w.release(); gd_cache.release(); pgd_buf.release(); prop_buf.release(); pos_w.release(); pub_rows.release();
        eta.release(); tau.release(); lik.release(); prior.release(); last4.release(); init_rmse.release();
        pub_lhood.release(); lik_prop.release(); rmse_tr.release(); rmse_te.release(); acc_tr.release();
        acc_te.release(); dbg_prior.release(); dbg_diff.release(); dbg_mh.release(); n_acc.release();
        init_count.release(); gd_valid.release(); accept_list.release(); dbg_acc.release(); swap_log.release();
        for (int q = 0; q < kMaxPeers; ++q) {
            if (peer_opened[q][0]) cudaIpcCloseMemHandle(peer_lhood[q]);
            if (peer_opened[q][1]) cudaIpcCloseMemHandle(peer_rows[q]);
            if (peer_opened[q][2]) cudaIpcCloseMemHandle(peer_flag_ptr[q]);
            peer_opened[q][0] = peer_opened[q][1] = peer_opened[q][2] = false;
        }
        peer_flags.release(); spec_bar.release(); spec_flag.release(); probe_count.release();
        barrier.release(); swap_counters.release(); d_lx.release(); d_z.release(); d_zeta.release();
        d_u.release(); d_uswap.release(); d_src.release(); smsp_load.release(); swap_src.release(); d_swapped.release(); d_scratch.release(); d_summary.release();
    }
};

static int fail(ptfnn_sampler *s, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (s) s->err = buf;
    return code;
}

// host mirrors of the cadence rules (R:427 | C:438, R:719)
static bool h_swap_due(int rule, int s, int i) { return rule == 0 ? (i % s == 0 && i != 0) : ((i + 1) % s == 0); }
static int h_inloop_rounds(const ptfnn_sampler *s, int upto_step_exclusive) {
    int n = 0;
    for (int i = 0; i < upto_step_exclusive; ++i) n += h_swap_due(s->swap_rule, s->cfg.swap_interval, i);
    return n;
}
static int h_total_rounds(const ptfnn_sampler *s) {
    const int n_r = h_inloop_rounds(s, s->cfg.samples - 1);
    const int n_m = s->cfg.samples / s->cfg.swap_interval;
    return n_r + (n_m > n_r ? 1 : 0);   // at most one left-over round on the exit vectors
}

// ------------------------------------------------------------------------------------------
// library-level entry points
// ------------------------------------------------------------------------------------------
extern "C" int ptfnn_abi_version(void) { return PTFNN_ABI_VERSION; }

extern "C" const char *ptfnn_build_info(void) {
    // one string per calling thread: the pointer stays valid for the caller while other threads register topologies
    static thread_local std::string info;
    char buf[256];
    snprintf(buf, sizeof buf, "libptfnn abi %d, sm_100a, nvcc %d.%d, tile_rows %d, topologies: ", PTFNN_ABI_VERSION,
             __CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__, kTileRows);
    info = buf + supported_list();
    return info.c_str();
}

// A further specialisation, compiled on demand from csrc/topo_inst.cu into its own shared library
// (capi.ensure_topology): both libraries link the SHARED CUDA runtime, so kernels of one can be launched
// by the other.  `kernel_set` points at that library's static PtfnnKernelSet (csrc/ptfnn_registry.h).
extern "C" int ptfnn_register_kernels(const void *kernel_set, int32_t registry_version) {
    if (!kernel_set) return fail(nullptr, PTFNN_E_INVALID, "null kernel set");
    if (registry_version != PTFNN_REGISTRY_VERSION) return fail(nullptr, PTFNN_E_INVALID, "kernel registry version %d != %d: rebuild the specialisation", registry_version, PTFNN_REGISTRY_VERSION);
    const KernelSet *k = (const KernelSet *)kernel_set;
    if (find_kernels(k->task, k->I, k->H, k->O)) return PTFNN_OK;
    std::lock_guard<std::mutex> g(kernel_sets_mutex());
    kernel_sets().push_back(k);
    return PTFNN_OK;
}

extern "C" int ptfnn_has_topology(int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out) {
    return find_kernels(task, n_in, n_hidden, n_out) ? 1 : 0;
}

extern "C" int ptfnn_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); cudaGetLastError(); return -1; }
    return n;
}

extern "C" void ptfnn_default_config(ptfnn_config *c) {
    memset(c, 0, sizeof *c);
    c->abi_version = PTFNN_ABI_VERSION;
    c->task = PTFNN_TASK_REGRESSION;
    c->swap_rule = PTFNN_SWAP_RULE_AUTO;
    c->use_langevin_gradients = 1;
    c->common_random_numbers = 1;
    c->memoize_gradient = 1;
    c->l_prob = 0.5;
    c->learn_rate = 0.1;
    c->step_w = 0.025;        // R:258
    c->step_eta = 0.2;        // R:260
    c->sigma_squared = 25.0;  // R:273
    c->nu_1 = 0.0; c->nu_2 = 0.0;
    c->pt_fraction = 0.6;     // R:301
    c->barrier_timeout_ms = 0;
    c->swap_kind = PTFNN_SWAP_KIND_REFERENCE;
}

extern "C" const char *ptfnn_last_error(const ptfnn_sampler *s) { return s ? s->err.c_str() : g_last_error.c_str(); }

static int require_device(ptfnn_sampler *s, int device) {
    const int n = ptfnn_device_count();
    if (n <= 0) return fail(s, PTFNN_E_CUDA, "no CUDA device available: libptfnn has no CPU path");
    if (device < 0 || device >= n) return fail(s, PTFNN_E_INVALID, "device %d out of range (have %d)", device, n);
    CU_TRY(s, cudaSetDevice(device));
    return PTFNN_OK;
}

// ------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------
extern "C" int ptfnn_create(const ptfnn_config *cfg, const double *temperatures, ptfnn_sampler **out) {
    if (!cfg || !temperatures || !out) return fail(nullptr, PTFNN_E_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != PTFNN_ABI_VERSION) return fail(nullptr, PTFNN_E_INVALID, "abi_version %d != %d", cfg->abi_version, PTFNN_ABI_VERSION);
    if (cfg->task != PTFNN_TASK_REGRESSION && cfg->task != PTFNN_TASK_CLASSIFICATION) return fail(nullptr, PTFNN_E_INVALID, "bad task %d", cfg->task);
    if (cfg->n_in < 1 || cfg->n_hidden < 1 || cfg->n_out < 1) return fail(nullptr, PTFNN_E_INVALID, "bad topology [%d,%d,%d]", cfg->n_in, cfg->n_hidden, cfg->n_out);
    if (cfg->task == PTFNN_TASK_REGRESSION && cfg->n_out != 1)
        return fail(nullptr, PTFNN_E_UNSUPPORTED, "regression needs one output (the reference stores fx[i] = out, R:132)");
    if (cfg->n_replicas < 1 || cfg->samples < 2 || cfg->swap_interval < 1) return fail(nullptr, PTFNN_E_INVALID, "need n_replicas >= 1, samples >= 2, swap_interval >= 1");
    if (cfg->swap_rule != PTFNN_SWAP_RULE_AUTO && cfg->swap_rule != PTFNN_SWAP_RULE_AFTER_I && cfg->swap_rule != PTFNN_SWAP_RULE_BEFORE_I1)
        return fail(nullptr, PTFNN_E_INVALID, "bad swap_rule %d", cfg->swap_rule);
    if (cfg->swap_kind != PTFNN_SWAP_KIND_REFERENCE && cfg->swap_kind != PTFNN_SWAP_KIND_RATIO_TEMPERATURE)
        return fail(nullptr, PTFNN_E_INVALID, "bad swap_kind %d", cfg->swap_kind);
    if (cfg->window_plan < 0 || cfg->window_plan > 2) return fail(nullptr, PTFNN_E_INVALID, "bad window_plan %d", cfg->window_plan);
    if (cfg->barrier_timeout_ms < 0) return fail(nullptr, PTFNN_E_INVALID, "barrier_timeout_ms %d < 0", cfg->barrier_timeout_ms);
    if (!(cfg->l_prob >= 0.0 && cfg->l_prob <= 1.0) || !std::isfinite(cfg->learn_rate) || !(cfg->step_w >= 0.0) || !(cfg->step_eta >= 0.0) ||
        !(cfg->sigma_squared > 0.0) || !std::isfinite(cfg->nu_1) || !std::isfinite(cfg->nu_2) || !(cfg->pt_fraction >= 0.0))
        return fail(nullptr, PTFNN_E_INVALID, "need 0 <= l_prob <= 1, finite learn_rate / nu, step_w >= 0, step_eta >= 0, sigma_squared > 0, pt_fraction >= 0");
    for (int k = 0; k < cfg->n_replicas; ++k)
        if (!(temperatures[k] > 0.0) || !std::isfinite(temperatures[k]))
            return fail(nullptr, PTFNN_E_INVALID, "temperature %d = %g: the likelihood is divided by it (R:204), need a finite value > 0", k, temperatures[k]);
    const int Rg = cfg->n_replicas_global > 0 ? cfg->n_replicas_global : cfg->n_replicas;
    if (cfg->replica_offset < 0 || cfg->replica_offset + cfg->n_replicas > Rg) return fail(nullptr, PTFNN_E_INVALID, "replica_offset/n_replicas outside the ladder of %d", Rg);
    if (cfg->swap_kind != PTFNN_SWAP_KIND_REFERENCE && Rg != cfg->n_replicas)
        return fail(nullptr, PTFNN_E_UNSUPPORTED, "swap_kind %d (the drafts' temperature-aware rule) is a single-GPU option", cfg->swap_kind);
    const KernelSet *ks = find_kernels(cfg->task, cfg->n_in, cfg->n_hidden, cfg->n_out);
    if (!ks) return fail(nullptr, PTFNN_E_UNSUPPORTED, "no sm_100a specialisation for %s topology [%d,%d,%d]; built: %s", cfg->task == kTaskReg ? "regression" : "classification", cfg->n_in, cfg->n_hidden, cfg->n_out, supported_list().c_str());
    int rc = require_device(nullptr, cfg->device);
    if (rc) return rc;

    ptfnn_sampler *s = new (std::nothrow) ptfnn_sampler();
    if (!s) return fail(nullptr, PTFNN_E_NOMEM, "out of host memory");
    s->cfg = *cfg;
    s->cfg.n_replicas_global = Rg;
    s->ks = ks;
    s->P = cfg->n_in * cfg->n_hidden + cfg->n_hidden * cfg->n_out + cfg->n_hidden + cfg->n_out;
    s->IP = (cfg->n_in + 3) & ~3;
    s->swap_rule = cfg->swap_rule == PTFNN_SWAP_RULE_AUTO ? (cfg->task == PTFNN_TASK_REGRESSION ? 0 : 1) : cfg->swap_rule;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) { rc = fail(nullptr, PTFNN_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); delete s; return rc; }
    s->num_sms = prop.multiProcessorCount;
    s->regs_per_sm = prop.regsPerMultiprocessor; s->threads_per_sm = prop.maxThreadsPerMultiProcessor;
    s->smem_per_sm = prop.sharedMemPerMultiprocessor;
    s->clock_khz = prop.clockRate;
    if (prop.major != 10) {                                                   // the kernels are sm_100a code only
        rc = fail(nullptr, PTFNN_E_UNSUPPORTED, "device %d (%s) is sm_%d%d: libptfnn is built for sm_100a (B200) only", cfg->device, prop.name, prop.major, prop.minor);
        delete s; return rc;
    }
    if (!prop.cooperativeLaunch) { rc = fail(nullptr, PTFNN_E_CUDA, "device lacks cooperative launch"); delete s; return rc; }

    const size_t R = cfg->n_replicas, S = cfg->samples, P = s->P;
    const int rounds = h_total_rounds(s) + 1;
#define ALLOC(buf, count)                                                                     \
    if ((e = s->buf.ensure(count)) != cudaSuccess) {                                          \
        rc = fail(nullptr, PTFNN_E_NOMEM, "cudaMalloc(" #buf ", %zu elems): %s", (size_t)(count), cudaGetErrorString(e)); \
        s->release_all(); delete s; cudaGetLastError(); return rc;                            \
    }
    ALLOC(temperature, R); ALLOC(w, R * P); ALLOC(gd_cache, R * P); ALLOC(pgd_buf, cfg->n_hidden > 64 ? R * P : 1); ALLOC(prop_buf, ks->fwd_tc ? R * P : 1);
    ALLOC(pos_w, R * S * P + kSumTracePadFloats); ALLOC(pub_rows, 2 * R * (P + kRowTail)); ALLOC(pub_lhood, 2 * (size_t)Rg);
    ALLOC(eta, R); ALLOC(tau, R); ALLOC(lik, R); ALLOC(prior, R); ALLOC(last4, R * 4); ALLOC(init_rmse, R * 2);
    ALLOC(lik_prop, R * S); ALLOC(rmse_tr, R * S); ALLOC(rmse_te, R * S); ALLOC(acc_tr, R * S); ALLOC(acc_te, R * S);
    ALLOC(n_acc, R); ALLOC(init_count, R); ALLOC(gd_valid, R); ALLOC(accept_list, R * S);
    if (cfg->debug_traces) { ALLOC(dbg_prior, R * S); ALLOC(dbg_diff, R * S); ALLOC(dbg_mh, R * S); ALLOC(dbg_acc, R * S); }
    ALLOC(swap_log, (size_t)rounds * std::max(Rg - 1, 1)); ALLOC(barrier, 1); ALLOC(swap_counters, 2); ALLOC(peer_flags, 2 * kMaxPeers); ALLOC(probe_count, 2); ALLOC(spec_bar, R); ALLOC(spec_flag, R * kSpecWords);
    ALLOC(d_src, (size_t)Rg); ALLOC(smsp_load, (size_t)s->num_sms * 4 + 64); ALLOC(swap_src, R); ALLOC(d_swapped, (size_t)std::max(Rg - 1, 1)); ALLOC(d_scratch, 16);
#undef ALLOC
    cudaMemcpy(s->temperature.p, temperatures, R * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(s->barrier.p, 0, sizeof(GridBarrier));
    cudaMemset(s->peer_flags.p, 0, 2 * kMaxPeers * sizeof(unsigned int));
    cudaMemset(s->probe_count.p, 0, 2 * sizeof(unsigned int));
    cudaMemset(s->spec_bar.p, 0, R * sizeof(GridBarrier));
    cudaMemset(s->spec_flag.p, 0, R * kSpecWords * sizeof(unsigned int));
    cudaMemset(s->smsp_load.p, 0, s->smsp_load.n * sizeof(int));
    cudaMemset(s->swap_counters.p, 0, 2 * sizeof(long long));
    cudaMemset(s->swap_log.p, 0, s->swap_log.n);
    e = cudaGetLastError();
    if (e != cudaSuccess) { rc = fail(nullptr, PTFNN_E_CUDA, "create: %s", cudaGetErrorString(e)); s->release_all(); delete s; return rc; }
    *out = s;
    return PTFNN_OK;
}

extern "C" int ptfnn_destroy(ptfnn_sampler *s) {
    if (!s) return PTFNN_OK;
    cudaSetDevice(s->cfg.device);
    cudaStreamSynchronize(s->stream);
    s->release_all();
    delete s;
    return PTFNN_OK;
}

extern "C" int ptfnn_set_stream(ptfnn_sampler *s, void *cuda_stream) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    s->stream = (cudaStream_t)cuda_stream;
    return PTFNN_OK;
}

static int pack_a_tiles(ptfnn_sampler *s, const KernelSet *ks, const float *x, int rows, int IP, DevBuf<float> &tiles, cudaStream_t st);
static int sync_and_check(ptfnn_sampler *s);
static const char *kDeviceFailedMsg = "a device-side wait of an earlier launch timed out (swap-round grid barrier or peer flags: the "
                                      "temperatures of one launch were not co-resident, or a peer rank never arrived); chain state and "
                                      "traces were left untouched -- call ptfnn_init_chains to start over";

// row-major float64 [rows, n_cols] -> padded float32 X [rows][IP] + y [pad4(rows)]
static void pack_dataset(const double *data, int rows, int n_cols, int I, int IP, float *x, float *y) {
    memset(x, 0, (size_t)rows * IP * sizeof(float));
    memset(y, 0, (size_t)((rows + 3) & ~3) * sizeof(float));
    for (int r = 0; r < rows; ++r) {
        for (int i = 0; i < I; ++i) x[(size_t)r * IP + i] = (float)data[(size_t)r * n_cols + i];
        y[r] = (float)data[(size_t)r * n_cols + I];
    }
}

// Classification labels index the class probabilities (C:217 prob[i, int(y)]; C:73-75 one-hot target): the
// reference raises IndexError for a label >= n_out and NumPy silently wraps a negative one; the device
// would read past the outputs, so both are refused here.  -1: all labels are fine.
static int first_bad_label(const double *data, int rows, int n_cols, int I, int O) {
    for (int r = 0; r < rows; ++r) {
        const double y = data[(size_t)r * n_cols + I];
        if (!(y > -1.0 && y < (double)O)) return r;              // int(y) truncates towards zero, like the reference
    }
    return -1;
}

extern "C" int ptfnn_set_data(ptfnn_sampler *s, const double *train, int32_t n_train, const double *test,
                              int32_t n_test, int32_t n_cols) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    if (!train || !test || n_train < 1 || n_test < 1) return fail(s, PTFNN_E_INVALID, "empty dataset");
    if (n_cols < s->cfg.n_in + 1) return fail(s, PTFNN_E_INVALID, "n_cols %d < n_in + 1", n_cols);
    if (s->cfg.task == PTFNN_TASK_CLASSIFICATION) {
        int bad = first_bad_label(train, n_train, n_cols, s->cfg.n_in, s->cfg.n_out);
        if (bad >= 0) return fail(s, PTFNN_E_INVALID, "train row %d: label %g outside [0, %d)", bad, train[(size_t)bad * n_cols + s->cfg.n_in], s->cfg.n_out);
        bad = first_bad_label(test, n_test, n_cols, s->cfg.n_in, s->cfg.n_out);
        if (bad >= 0) return fail(s, PTFNN_E_INVALID, "test row %d: label %g outside [0, %d)", bad, test[(size_t)bad * n_cols + s->cfg.n_in], s->cfg.n_out);
    }
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    // The rows are packed (fp32, padded to IP floats) into one of two page-locked staging slots and copied behind
    // whatever the stream is doing: the call does not wait for the device -- a caller that uploads a data set between
    // two launches keeps the GPU busy meanwhile (bench.py's e2e loop).  A slot is reused two calls later, after the
    // event recorded behind its copies.
    const int IP = s->IP, I = s->cfg.n_in;
    const size_t nx_tr = (size_t)n_train * IP, ny_tr = (size_t)((n_train + 3) & ~3);
    const size_t nx_te = (size_t)n_test * IP, ny_te = (size_t)((n_test + 3) & ~3);
    const size_t need = (nx_tr + ny_tr + nx_te + ny_te) * sizeof(float);
    ptfnn_sampler::StageSlot &st = s->stage[s->stage_seq & 1];
    s->stage_seq += 1;
    if (st.used) CU_TRY(s, cudaEventSynchronize(st.done));
    if (st.bytes < need) {
        if (st.buf) CU_TRY(s, cudaFreeHost(st.buf));
        st.buf = nullptr; st.bytes = 0;
        CU_TRY(s, cudaHostAlloc((void **)&st.buf, need, cudaHostAllocDefault));
        st.bytes = need;
    }
    if (!st.done) CU_TRY(s, cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
    float *hx_tr = st.buf, *hy_tr = hx_tr + nx_tr, *hx_te = hy_tr + ny_tr, *hy_te = hx_te + nx_te;
    pack_dataset(train, n_train, n_cols, I, IP, hx_tr, hy_tr);
    pack_dataset(test, n_test, n_cols, I, IP, hx_te, hy_te);
    CU_TRY(s, s->train_x.ensure(nx_tr)); CU_TRY(s, s->train_y.ensure(ny_tr));
    CU_TRY(s, s->test_x.ensure(nx_te)); CU_TRY(s, s->test_y.ensure(ny_te));
    CU_TRY(s, cudaMemcpyAsync(s->train_x.p, hx_tr, nx_tr * 4, cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaMemcpyAsync(s->train_y.p, hy_tr, ny_tr * 4, cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaMemcpyAsync(s->test_x.p, hx_te, nx_te * 4, cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaMemcpyAsync(s->test_y.p, hy_te, ny_te * 4, cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaEventRecord(st.done, s->stream));
    st.used = true;
    s->n_train = n_train; s->n_test = n_test;
    if (s->ks->fwd_tc) {
        int rc = pack_a_tiles(s, s->ks, s->train_x.p, n_train, s->IP, s->a_train, s->stream);
        if (!rc) rc = pack_a_tiles(s, s->ks, s->test_x.p, n_test, s->IP, s->a_test, s->stream);
        if (rc) return rc;
    }
    s->have_data = true;
    return PTFNN_OK;
}

static DataView view(const DevBuf<float> &x, const DevBuf<float> &y, int n) { return DataView{x.p, y.p, n}; }

static int upload_doubles_as_float(ptfnn_sampler *s, float *dst, const double *src, size_t n) {
    std::vector<float> tmp(n);
    for (size_t i = 0; i < n; ++i) tmp[i] = (float)src[i];
    CU_TRY(s, cudaMemcpyAsync(dst, tmp.data(), n * 4, cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return PTFNN_OK;
}

extern "C" int ptfnn_init_chains(ptfnn_sampler *s, const double *w) {
    if (!s || !w) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!s->have_data) return fail(s, PTFNN_E_STATE, "ptfnn_set_data must come first");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const size_t R = s->cfg.n_replicas, S = s->cfg.samples, P = s->P;
    int rc = upload_doubles_as_float(s, s->w.p, w, R * P);
    if (rc) return rc;
    // trace row 0 and carried rows: pos_w = 1, likelihood = -100, everything else 0
    {
        std::vector<float> ones(R * P, 1.0f);
        for (size_t r = 0; r < R; ++r)
            CU_TRY(s, cudaMemcpyAsync(s->pos_w.p + r * S * P, ones.data(), P * 4, cudaMemcpyHostToDevice, s->stream));
        std::vector<double> m100(R * S, 0.0);
        for (size_t r = 0; r < R; ++r) m100[r * S] = -100.0;     // R:293
        CU_TRY(s, cudaMemcpyAsync(s->lik_prop.p, m100.data(), R * S * 8, cudaMemcpyHostToDevice, s->stream));
        CU_TRY(s, cudaStreamSynchronize(s->stream));
    }
    CU_TRY(s, cudaMemsetAsync(s->rmse_tr.p, 0, R * S * 8, s->stream)); CU_TRY(s, cudaMemsetAsync(s->rmse_te.p, 0, R * S * 8, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->acc_tr.p, 0, R * S * 8, s->stream)); CU_TRY(s, cudaMemsetAsync(s->acc_te.p, 0, R * S * 8, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->accept_list.p, 0, R * S * 4, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->last4.p, 0, R * 4 * 8, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->n_acc.p, 0, R * 4, s->stream)); CU_TRY(s, cudaMemsetAsync(s->init_count.p, 0, R * 4, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->gd_valid.p, 0, R * 4, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->swap_counters.p, 0, 16, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->swap_log.p, 0, s->swap_log.n, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->barrier.p, 0, sizeof(GridBarrier), s->stream));
    CU_TRY(s, cudaMemsetAsync(s->spec_bar.p, 0, R * sizeof(GridBarrier), s->stream));
    CU_TRY(s, cudaMemsetAsync(s->spec_flag.p, 0, R * kSpecWords * sizeof(unsigned int), s->stream));
    if (s->cfg.debug_traces) {
        CU_TRY(s, cudaMemsetAsync(s->dbg_prior.p, 0, R * S * 8, s->stream)); CU_TRY(s, cudaMemsetAsync(s->dbg_diff.p, 0, R * S * 8, s->stream));
        CU_TRY(s, cudaMemsetAsync(s->dbg_mh.p, 0, R * S * 8, s->stream)); CU_TRY(s, cudaMemsetAsync(s->dbg_acc.p, 0, R * S, s->stream));
    }
    InitParams ip;
    ip.R = (int)R; ip.S = (int)S;
    ip.sigma_sq = s->cfg.sigma_squared; ip.nu1 = s->cfg.nu_1; ip.nu2 = s->cfg.nu_2;
    ip.temperature = s->temperature.p;
    ip.train = view(s->train_x, s->train_y, s->n_train);
    ip.test = view(s->test_x, s->test_y, s->n_test);
    ip.w = s->w.p; ip.eta = s->eta.p; ip.tau = s->tau.p; ip.lik = s->lik.p; ip.prior = s->prior.p;
    ip.init_rmse = s->init_rmse.p;
    void *args[] = {&ip};
    const size_t smem = ((P * 4 + 15) & ~(size_t)15) + 8 * (s->ks->NT / 32) * 8 + 64;
    CU_TRY(s, cudaFuncSetAttribute(s->ks->init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_TRY(s, cudaLaunchKernel(s->ks->init, dim3((unsigned)R), dim3(s->ks->NT), args, smem, s->stream));
    CU_TRY(s, cudaGetLastError());
    // The peer arrival flags are written remotely and never reset: a re-initialised handle moves to a fresh
    // range of the flag domain instead (same arithmetic on every rank), so that the counts of the previous
    // run are not taken for "already published".
    if (s->have_state) s->peer_round_base += (unsigned int)h_total_rounds(s) + 2u;
    s->device_failed = false;
    s->acc_est = -1.0; s->acc_have_prev = false; s->acc_prev_sum = 0; s->acc_prev_step = 0;
    for (auto &a : s->acc_ring) a.used = false;
    s->arm_cost[0] = s->arm_cost[1] = -1.0; s->arm_launched[0] = s->arm_launched[1] = -1; s->arm_acc[0] = s->arm_acc[1] = 1.0; s->arm_timed[0] = s->arm_timed[1] = -100; s->arm_interval[0] = s->arm_interval[1] = 32; s->fb_launches = 0;
    s->step = 0; s->rounds_done = 0; s->swap_pending = false; s->pending_final = false;
    s->host_num_swap = 0; s->host_total_prop = 0; s->host_swap_log.clear(); s->host_swap_log_round.clear();
    s->have_state = true;
    return PTFNN_OK;
}

extern "C" int ptfnn_set_state(ptfnn_sampler *s, const double *w, const double *eta, const double *lik,
                               const double *prior, const double *tau) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    if (!s->have_state) return fail(s, PTFNN_E_STATE, "ptfnn_init_chains must come first");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const size_t R = s->cfg.n_replicas, P = s->P;
    if (w) {
        int rc = upload_doubles_as_float(s, s->w.p, w, R * P);
        if (rc) return rc;
        CU_TRY(s, cudaMemsetAsync(s->gd_valid.p, 0, R * 4, s->stream));
    }
    if (eta) CU_TRY(s, cudaMemcpyAsync(s->eta.p, eta, R * 8, cudaMemcpyHostToDevice, s->stream));
    if (lik) CU_TRY(s, cudaMemcpyAsync(s->lik.p, lik, R * 8, cudaMemcpyHostToDevice, s->stream));
    if (prior) CU_TRY(s, cudaMemcpyAsync(s->prior.p, prior, R * 8, cudaMemcpyHostToDevice, s->stream));
    if (tau) CU_TRY(s, cudaMemcpyAsync(s->tau.p, tau, R * 8, cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return PTFNN_OK;
}

extern "C" int ptfnn_get_state(ptfnn_sampler *s, double *w, double *eta, double *lik, double *prior, double *tau,
                               int32_t *num_accepted) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    if (!s->have_state) return fail(s, PTFNN_E_STATE, "ptfnn_init_chains must come first");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const size_t R = s->cfg.n_replicas, P = s->P;
    { const int rc = sync_and_check(s); if (rc) return rc; }
    if (w) {
        std::vector<float> tmp(R * P);
        CU_TRY(s, cudaMemcpy(tmp.data(), s->w.p, R * P * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < R * P; ++i) w[i] = tmp[i];
    }
    if (eta) CU_TRY(s, cudaMemcpy(eta, s->eta.p, R * 8, cudaMemcpyDeviceToHost));
    if (lik) CU_TRY(s, cudaMemcpy(lik, s->lik.p, R * 8, cudaMemcpyDeviceToHost));
    if (prior) CU_TRY(s, cudaMemcpy(prior, s->prior.p, R * 8, cudaMemcpyDeviceToHost));
    if (tau) CU_TRY(s, cudaMemcpy(tau, s->tau.p, R * 8, cudaMemcpyDeviceToHost));
    if (num_accepted) CU_TRY(s, cudaMemcpy(num_accepted, s->n_acc.p, R * 4, cudaMemcpyDeviceToHost));
    return PTFNN_OK;
}

extern "C" int ptfnn_get_step(const ptfnn_sampler *s, int32_t *step, int32_t *swap_rounds_done) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    if (step) *step = s->step;
    if (swap_rounds_done) *swap_rounds_done = s->rounds_done;
    return PTFNN_OK;
}

// ------------------------------------------------------------------------------------------
// the hot path
// ------------------------------------------------------------------------------------------
static const size_t kStageLimitBytes = 96 * 1024;

static int launch_chain(ptfnn_sampler *s, int n_steps, const ptfnn_draws *d, int32_t *steps_done) {
    if (steps_done) *steps_done = 0;
    if (!s->have_state) return fail(s, PTFNN_E_STATE, "ptfnn_init_chains must come first");
    if (s->swap_pending) return fail(s, PTFNN_E_STATE, "a swap round is pending: finish it with ptfnn_swap_plan/apply");
    if (s->device_failed) return fail(s, PTFNN_E_CUDA, "%s", kDeviceFailedMsg);
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const ptfnn_config &c = s->cfg;
    const int R = c.n_replicas, Rg = c.n_replicas_global, S = c.samples, P = s->P;
    const bool external = Rg > R && s->n_ranks == 1;   // host-completed rounds unless the ranks are peer-connected
    int n = std::min(n_steps, S - 1 - s->step);
    if (n <= 0) return PTFNN_OK;
    // multi-GPU: stop right after the first step at which a swap is due
    if (external)
        for (int k = 0; k < n; ++k)
            if (h_swap_due(s->swap_rule, c.swap_interval, s->step + k)) { n = k + 1; break; }
    const int begin = s->step, end = s->step + n;
    int rounds_in_span = 0;
    if (Rg > 1)
        for (int i = begin; i < end; ++i) rounds_in_span += h_swap_due(s->swap_rule, c.swap_interval, i);
    const bool chain_done = end == S - 1;
    const bool final_round = Rg > 1 && chain_done && (h_total_rounds(s) > h_inloop_rounds(s, S - 1));

    ChainParams p;
    memset(&p, 0, sizeof p);
    p.R = R; p.Rg = Rg; p.replica_offset = c.replica_offset;
    p.S = S; p.swap_interval = c.swap_interval; p.swap_rule = s->swap_rule;
    p.use_lg = c.use_langevin_gradients; p.crn = c.common_random_numbers; p.memo = c.memoize_gradient;
    p.external_swap = external; p.debug = c.debug_traces;
    p.step_begin = begin; p.step_end = end; p.round_begin = s->rounds_done;
    p.final_round = final_round && !external;
    p.l_prob = c.l_prob; p.pt_samples = (double)S * c.pt_fraction;   // R:301 `samples * 0.6` (a float)
    p.sigma_sq = c.sigma_squared; p.nu1 = c.nu_1; p.nu2 = c.nu_2;
    p.lr = (float)c.learn_rate; p.step_w = (float)c.step_w; p.step_eta = (float)c.step_eta;
    p.seed = c.seed;
    p.temperature = s->temperature.p;
    p.train = view(s->train_x, s->train_y, s->n_train);
    p.test = view(s->test_x, s->test_y, s->n_test);
    p.a_train = s->a_train.p; p.a_test = s->a_test.p;
    p.w = s->w.p; p.eta = s->eta.p; p.tau = s->tau.p; p.lik = s->lik.p; p.prior = s->prior.p;
    p.n_acc = s->n_acc.p; p.init_count = s->init_count.p; p.last4 = s->last4.p;
    p.gd_cache = s->gd_cache.p; p.pgd_buf = s->pgd_buf.p; p.prop_buf = s->prop_buf.p; p.gd_valid = s->gd_valid.p;
    p.pos_w = s->pos_w.p; p.lik_prop = s->lik_prop.p; p.rmse_tr = s->rmse_tr.p; p.rmse_te = s->rmse_te.p;
    p.acc_tr = s->acc_tr.p; p.acc_te = s->acc_te.p; p.accept_list = s->accept_list.p;
    p.dbg_prior = s->dbg_prior.p; p.dbg_diff = s->dbg_diff.p; p.dbg_mh = s->dbg_mh.p; p.dbg_acc = s->dbg_acc.p;
    p.pub_rows = s->pub_rows.p; p.pub_lhood = s->pub_lhood.p; p.barrier = s->barrier.p;
    p.swap_counters = s->swap_counters.p; p.swap_log = s->swap_log.p;
    p.max_rounds = (int)(s->swap_log.n / std::max(Rg - 1, 1));
    p.smsp_load = s->smsp_load.p;
    p.n_ranks = s->n_ranks; p.rank = s->rank;
    for (int q = 0; q < kMaxPeers; ++q) {
        p.peer_lhood[q] = (double *)s->peer_lhood[q]; p.peer_rows[q] = (const float *)s->peer_rows[q];
        p.peer_flags[q] = (unsigned int *)s->peer_flag_ptr[q];
    }
    p.swap_src = s->swap_src.p;
    p.P = P;
    p.peer_round_base = s->peer_round_base;
    p.heartbeat = s->peer_flags.p + kMaxPeers;
    p.wait_limit = (long long)(c.barrier_timeout_ms > 0 ? c.barrier_timeout_ms : 20000) * (long long)std::max(s->clock_khz, 1);
    p.probe = 0; p.probe_count = s->probe_count.p;
    p.swap_kind = c.swap_kind; p.temperature_global = s->temperature.p;   // swap_kind != 0 is single-GPU: the local ladder is the ladder

    if (d) {
        if (d->n < n) return fail(s, PTFNN_E_INVALID, "draws cover %d steps, need %d", d->n, n);
        if (!d->lx || !d->z || !d->u) return fail(s, PTFNN_E_INVALID, "draws: lx, z and u are required");
        const int need_rounds = external ? 0 : rounds_in_span + (p.final_round ? 1 : 0);
        if (need_rounds > 0 && (!d->u_swap || d->n_swap_rounds < need_rounds))
            return fail(s, PTFNN_E_INVALID, "draws: need %d swap rounds of uniforms, got %d", need_rounds, d->u_swap ? d->n_swap_rounds : 0);
        const size_t rn = (size_t)R * d->n;
        CU_TRY(s, s->d_lx.ensure(rn)); CU_TRY(s, s->d_u.ensure(rn)); CU_TRY(s, s->d_z.ensure(rn * P)); CU_TRY(s, s->d_zeta.ensure(rn));
        CU_TRY(s, cudaMemcpyAsync(s->d_lx.p, d->lx, rn * 4, cudaMemcpyHostToDevice, s->stream));
        CU_TRY(s, cudaMemcpyAsync(s->d_u.p, d->u, rn * 4, cudaMemcpyHostToDevice, s->stream));
        CU_TRY(s, cudaMemcpyAsync(s->d_z.p, d->z, rn * P * 4, cudaMemcpyHostToDevice, s->stream));
        if (d->z_eta) CU_TRY(s, cudaMemcpyAsync(s->d_zeta.p, d->z_eta, rn * 4, cudaMemcpyHostToDevice, s->stream));
        if (need_rounds > 0) {
            const size_t nu = (size_t)need_rounds * (Rg - 1);
            CU_TRY(s, s->d_uswap.ensure(nu));
            CU_TRY(s, cudaMemcpyAsync(s->d_uswap.p, d->u_swap, nu * 4, cudaMemcpyHostToDevice, s->stream));
        }
        p.replay = 1; p.replay_n = d->n;
        p.lx = s->d_lx.p; p.z = s->d_z.p; p.z_eta = d->z_eta ? s->d_zeta.p : nullptr; p.u = s->d_u.p; p.u_swap = s->d_uswap.p;
    }

    const size_t stage_bytes = ((size_t)s->n_train * s->IP + ((s->n_train + 3) & ~3) + (size_t)s->n_test * s->IP + ((s->n_test + 3) & ~3)) * 4;
    p.staged = stage_bytes <= kStageLimitBytes ? 1 : 0;
    const int NT = s->ks->NT;
    const int team_floats = c.n_hidden > 64 ? team_smem_floats(c.n_hidden, c.n_out) : 0;   // mirrors UseSgdTeam<H>
    const int lik_floats = c.n_hidden * ((c.n_in + 1 + c.n_out + 3) & ~3);   // mirrors LikLayout<I, O>::LW
    const ChainSmem L = chain_smem_layout(P, s->IP, NT, external ? 1 : Rg, p.staged != 0, s->n_train, s->n_test, team_floats, lik_floats, team_floats == 0,
                                          s->ks->fwd_tc ? s->ks->tc_smem_bytes : 0, s->ks->tc_alias_off);
    if (L.total > 227 * 1024) return fail(s, PTFNN_E_UNSUPPORTED, "needs %zu bytes of shared memory per CTA (> 227 KB)", L.total);
    CU_TRY(s, cudaFuncSetAttribute(s->ks->chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    CU_TRY(s, cudaFuncSetAttribute(s->ks->chain, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    CU_TRY(s, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s->ks->chain, NT, L.total));
    if (per_sm < 1) return fail(s, PTFNN_E_CUDA, "chain kernel does not fit on an SM (smem %zu)", L.total);
    const bool uses_tmem = s->ks->fwd_tc != nullptr;
    if (uses_tmem) {
        // The occupancy calculator answers 1 for every kernel that executes tcgen05.alloc, although the
        // hardware co-schedules CTAs as long as their TMEM columns fit (tools/occ_probe.cu: two CTAs of
        // 256 columns per SM), so such kernels cannot be launched cooperatively.  The estimate below
        // (registers / shared memory / threads / TMEM columns) is only the starting point: co-residency is
        // MEASURED by a probe launch of the real kernel with the real launch configuration, and the grid is
        // clamped to the number of CTAs that were resident together (other contexts on the GPU, driver-
        // reserved shared memory and allocation granularities are then accounted for by construction).
        cudaFuncAttributes fa;
        CU_TRY(s, cudaFuncGetAttributes(&fa, s->ks->chain));
        const int regs = ((fa.numRegs + 7) / 8) * 8;
        const int by_regs = s->regs_per_sm / std::max(1, regs * NT);
        const int by_smem = (int)(s->smem_per_sm / (L.total + fa.sharedSizeBytes + 1024));
        const int by_thr = s->threads_per_sm / NT;
        const int by_tmem = 512 / std::max(32, s->ks->tmem_cols);
        per_sm = std::max(1, std::min(std::min(by_regs, by_smem), std::min(by_thr, by_tmem)));
        const int want = std::min(R, per_sm * s->num_sms);
        if (s->verified_grid < 1 || s->verified_smem != L.total || s->verified_grid > want) {
            int try_grid = want, proven = 0;
            for (int attempt = 0; attempt < 4 && try_grid >= 1; ++attempt) {
                CU_TRY(s, cudaMemsetAsync(s->probe_count.p, 0, 2 * sizeof(unsigned int), s->stream));
                ChainParams pp = p;
                pp.probe = 1;
                void *pargs[] = {&pp};
                CU_TRY(s, cudaLaunchKernel(s->ks->chain, dim3(try_grid), dim3(NT), pargs, L.total, s->stream));
                unsigned int seen[2] = {0, 0};
                CU_TRY(s, cudaMemcpyAsync(seen, s->probe_count.p, sizeof seen, cudaMemcpyDeviceToHost, s->stream));
                CU_TRY(s, cudaStreamSynchronize(s->stream));
                if ((int)seen[1] >= try_grid) { proven = try_grid; break; }
                // fewer CTAs than asked were resident together: whole CTAs per SM, one less than the estimate
                try_grid = std::min((int)seen[1], (try_grid > s->num_sms ? (try_grid - 1) / s->num_sms : 0) * s->num_sms);
                if (try_grid < 1) try_grid = std::min((int)seen[1], s->num_sms);
            }
            if (proven < 1) return fail(s, PTFNN_E_CUDA, "the chain kernel's CTAs are not co-resident on device %d even at one per SM (another context holding the GPU?)", c.device);
            s->verified_grid = proven; s->verified_smem = L.total;
        }
        per_sm = std::max(1, s->verified_grid / s->num_sms);
    }
    int grid = uses_tmem ? std::min(R, s->verified_grid) : std::min(R, per_sm * s->num_sms);
    // Small ladders leave most of the GPU idle (10 temperatures = 10 of 148 SMs): K CTAs per temperature
    // evaluate the steps ahead speculatively, a Langevin step and the random-walk steps before it per CTA
    // (chain_body, "speculative windows").  cfg.speculation: 0 = automatic, 1 = off, K > 1 = that many CTAs
    // (clamped to what is co-resident).
    int spec = 1;
    const bool spec_possible = s->ks->chain_spec && !external && R * 2 <= per_sm * s->num_sms;
    const bool spec_auto = spec_possible && c.speculation == 0;
    const bool spec_feedback = spec_auto && R * 2 > s->num_sms;    // (tiny ladders do not steer by the acceptance rate)
    if (spec_feedback) {
        if (s->acc_seq >= 2) {                           // bound the lag: the sample of the launch before the previous one
            ptfnn_sampler::AccSample &old = s->acc_ring[(s->acc_seq - 2) & 3];
            if (old.used) CU_TRY(s, cudaEventSynchronize(old.ev));
        }
        for (int k = 3; k >= 0; --k) {                   // launch timings, oldest first, each sample once
            ptfnn_sampler::AccSample &a = s->acc_ring[(s->acc_seq + 3 - k) & 3];
            if (!a.used || a.timed || cudaEventQuery(a.ev) != cudaSuccess) continue;
            a.timed = true;
            float ms = 0.0f;
            const double units = (double)kSpecRwRun * (double)a.host[1] + (double)a.host[2];
            if (cudaEventElapsedTime(&ms, a.t0, a.t1) == cudaSuccess && units > 0.0) {
                const double c = (double)ms / units;
                // (a timing after a pause replaces the old estimate: the run has moved on meanwhile)
                const bool retry = a.launch - s->arm_timed[a.arm] > 4;
                s->arm_cost[a.arm] = (s->arm_cost[a.arm] < 0.0 || retry) ? c : 0.5 * s->arm_cost[a.arm] + 0.5 * c;
                s->arm_timed[a.arm] = a.launch;
                if (retry && s->arm_cost[1 - a.arm] >= 0.0)       // a retry that lost again waits twice as long for the next one
                    s->arm_interval[a.arm] = s->arm_cost[a.arm] > s->arm_cost[1 - a.arm] ? std::min(512, 2 * s->arm_interval[a.arm]) : 32;
            }
        }
        for (int k = 0; k < 4; ++k) {
            ptfnn_sampler::AccSample &a = s->acc_ring[(s->acc_seq + 3 - k) & 3];
            if (!a.used || cudaEventQuery(a.ev) != cudaSuccess) continue;
            if (s->acc_have_prev && a.step > s->acc_prev_step) {
                const double rate = (double)(a.host[0] - s->acc_prev_sum) / ((double)R * (a.step - s->acc_prev_step));
                s->acc_est = s->acc_est < 0.0 ? rate : 0.5 * s->acc_est + 0.5 * rate;
            }
            if (!s->acc_have_prev || a.step > s->acc_prev_step) { s->acc_prev_sum = a.host[0]; s->acc_prev_step = a.step; s->acc_have_prev = true; }
            break;
        }
    }
    if (spec_possible) {
        // A window of depth K costs about one step (of the slowest kind in it) and advances (1 - (1-a)^K) / a steps
        // for acceptance rate a: K pays as soon as proposals are mostly rejected, which is the reference's steady
        // state (12-30 %; a few per cent on 30 000-row data sets) but not the first steps of a run from
        // random weights, where Langevin proposals are accepted almost always and extra CTAs only add contention
        // (measured in round 1: 42-48 ms against 36-40 ms per 10 steps of 128-512 temperatures at acceptance ~1).
        //   automatic:  ladder <= half the SMs (the reference's own 10 temperatures): one CTA per SM, as before;
        //               otherwise the depth follows the acceptance rate observed on THIS run so far.
        const int cap = std::min(kMaxSpec, per_sm * s->num_sms / R);
        int want = c.speculation;
        if (want == 0) {
            const double a = s->acc_est;
            if (R * 2 <= s->num_sms)      // tiny ladders: one CTA per SM for Langevin runs; none for random-walk runs (a step is shorter than a window's two group barriers)
                want = c.use_langevin_gradients ? std::max(1, s->num_sms / R) : 1;
            else
                // (measured, 4-64-1: two CTAs per temperature pay below ~4 % acceptance -- the doubled residency slows every
                //  step by a quarter -- four or more already at 8 %: 31.5 against 38.8 ms per 10 steps of 128 temperatures)
                want = a < 0.0 ? 1 : cap <= 2 ? (a <= 0.04 ? cap : 1) : cap == 3 ? (a <= 0.08 ? cap : 1) : a <= 0.30 ? cap : a <= 0.60 ? std::min(cap, 4) : 1;
            // ... and the clock has the last word.  The rate above is this rank's mean; what a segment costs with windows
            // is set by the temperature with the MOST acceptances in it, of the whole ladder (every one of them costs that
            // temperature another window and everybody meets at the swap round): 8 GPUs x 128 temperatures at 11 % ran
            // 45.5 ms per 10 steps with windows against ~36 sequentially, although 128 temperatures alone gain.  So both
            // ways of running are timed on this run (cost per random-walk-equivalent step, see the harvest above), the
            // cheaper one is used, and the other one is tried again every 32 launches -- twice as many after every try it loses.
            if (spec_feedback && want > 1) {                       // (where the rate rules windows out they are not tried either)
                const int L = s->fb_launches;
                const int kwin = want;
                // (stale: not run for kReprobe launches, or the acceptance rate has fallen by a third since -- windows get
                //  cheaper quickly while a run burns in)
                auto stale = [&](int arm) { return s->arm_launched[arm] < 0 || L - s->arm_launched[arm] >= s->arm_interval[arm] || (arm == 1 && a < 0.67 * s->arm_acc[1]); };
                int arm;
                if (stale(0)) arm = 0;
                else if (stale(1)) arm = 1;
                else if (s->arm_cost[0] < 0.0) arm = 0;            // launched, not timed yet: stay until it is
                else if (s->arm_cost[1] < 0.0) arm = 1;
                else arm = s->arm_cost[1] < 0.97 * s->arm_cost[0] ? 1 : 0;
                want = arm ? kwin : 1;
            }
        }
        spec = std::max(1, std::min(want, cap));
    }
    p.spec_k = spec; p.spec_bar = s->spec_bar.p; p.spec_flag = s->spec_flag.p;
    p.spec_plan = c.window_plan;                       // measurement knob (tools/): results do not depend on it
    const void *chain_fn = s->ks->chain;
    if (spec > 1) {
        chain_fn = s->ks->chain_spec;
        CU_TRY(s, cudaFuncSetAttribute(chain_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        int per_sm_spec = 0;
        CU_TRY(s, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_spec, chain_fn, NT, L.total));
        spec = std::max(1, std::min(spec, per_sm_spec * s->num_sms / R));
        p.spec_k = spec;
        if (spec > 1) grid = R * spec; else chain_fn = s->ks->chain;
    }
    // many temperatures per SM: keep the serial warps' sub-partitions quiet (see chain_kernel)
    void *args[] = {&p};
    ptfnn_sampler::AccSample *fb = nullptr;
    if (spec_feedback) {
        fb = &s->acc_ring[s->acc_seq & 3];
        if (!fb->host) {
            CU_TRY(s, cudaHostAlloc((void **)&fb->host, 3 * sizeof(long long), cudaHostAllocDefault));
            CU_TRY(s, cudaEventCreateWithFlags(&fb->ev, cudaEventDisableTiming));
            CU_TRY(s, cudaEventCreate(&fb->t0));
            CU_TRY(s, cudaEventCreate(&fb->t1));
        }
        if (fb->used && cudaEventSynchronize(fb->ev) != cudaSuccess) fb = nullptr;   // (the slot of four launches ago: long finished)
    }
    if (fb) CU_TRY(s, cudaEventRecord(fb->t0, s->stream));
    if (uses_tmem) CU_TRY(s, cudaLaunchKernel(s->ks->chain, dim3(grid), dim3(NT), args, L.total, s->stream));
    else CU_TRY(s, cudaLaunchCooperativeKernel(chain_fn, dim3(grid), dim3(NT), args, L.total, s->stream));
    CU_TRY(s, cudaGetLastError());
    s->step = end;
    if (fb) {
        CU_TRY(s, cudaEventRecord(fb->t1, s->stream));
        CU_TRY(s, s->acc_sum.ensure(3));
        const uint32_t gr0 = (uint32_t)c.replica_offset;
        feedback_kernel<<<1, 256, 0, s->stream>>>(s->n_acc.p, R, c.seed, begin, n, c.common_random_numbers ? kStreamCommon : gr0, gr0,
                                                  d ? p.lx : nullptr, c.use_langevin_gradients ? 1 : 0, (double)c.l_prob, s->acc_sum.p);
        CU_TRY(s, cudaMemcpyAsync(fb->host, s->acc_sum.p, 3 * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
        CU_TRY(s, cudaEventRecord(fb->ev, s->stream));
        fb->step = end; fb->used = true; fb->timed = false; fb->arm = spec > 1 ? 1 : 0; fb->launch = s->fb_launches;
        s->arm_launched[fb->arm] = s->fb_launches; s->arm_acc[fb->arm] = s->acc_est < 0.0 ? 1.0 : s->acc_est;
        s->fb_launches += 1;
        s->acc_seq += 1;
    }
    if (external) {
        if (rounds_in_span > 0) { s->swap_pending = true; s->pending_final = false; }
        else if (final_round) { s->swap_pending = true; s->pending_final = true; }
    } else {
        s->rounds_done += rounds_in_span + (p.final_round ? 1 : 0);
    }
    if (steps_done) *steps_done = n;
    return PTFNN_OK;
}

extern "C" int ptfnn_run(ptfnn_sampler *s, int32_t n_steps, int32_t *steps_done) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    // Ladders on which the automatic speculation steers by what earlier launches of the run showed (launch_chain:
    // acceptance rate, launch times) get a long request in pieces that end on swap rounds -- otherwise a caller that
    // asks for the whole chain at once would run all of it on the decision of step 0.  Pieces cost a launch (~10 us) per
    // >= 64 steps.  (Replayed draws, fixed depths, tiny and full ladders: one launch, as asked.)
    const ptfnn_config &c = s->cfg;
    const bool host_rounds = c.n_replicas_global > c.n_replicas && s->n_ranks == 1;
    const bool adaptive = s->ks && s->ks->chain_spec && c.speculation == 0 && !host_rounds && c.n_replicas_global > 1 &&
                          2 * c.n_replicas > s->num_sms && c.n_replicas <= 8 * s->num_sms;
    if (!adaptive) return launch_chain(s, n_steps, nullptr, steps_done);
    if (steps_done) *steps_done = 0;
    const int rounds_per_piece = std::max(1, (64 + c.swap_interval - 1) / c.swap_interval);
    int remaining = n_steps, total = 0;
    while (remaining > 0) {
        int piece = remaining, rounds = 0;
        for (int k = 0; k < remaining; ++k)
            if (h_swap_due(s->swap_rule, c.swap_interval, s->step + k) && ++rounds == rounds_per_piece) { piece = k + 1; break; }
        int32_t done = 0;
        const int rc = launch_chain(s, piece, nullptr, &done);
        if (rc != PTFNN_OK) return rc;
        total += done; remaining -= piece;
        if (steps_done) *steps_done = total;
        if (done < piece) break;          // the end of the chain
    }
    return PTFNN_OK;
}

extern "C" int ptfnn_replay(ptfnn_sampler *s, const ptfnn_draws *d, int32_t *steps_done) {
    if (!s || !d) return fail(s, PTFNN_E_INVALID, "null argument");
    return launch_chain(s, d->n, d, steps_done);
}

// stream synchronisation + the device-side failure flag of the grid barrier
static int sync_and_check(ptfnn_sampler *s) {
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    if (s->device_failed) return fail(s, PTFNN_E_CUDA, "%s", kDeviceFailedMsg);
    GridBarrier b;
    CU_TRY(s, cudaMemcpy(&b, s->barrier.p, sizeof b, cudaMemcpyDeviceToHost));
    std::vector<GridBarrier> sb(s->cfg.n_replicas);
    CU_TRY(s, cudaMemcpy(sb.data(), s->spec_bar.p, sb.size() * sizeof(GridBarrier), cudaMemcpyDeviceToHost));
    for (const GridBarrier &g : sb) b.failed |= g.failed;
    if (b.failed) { s->device_failed = true; return fail(s, PTFNN_E_CUDA, "%s", kDeviceFailedMsg); }
    return PTFNN_OK;
}

extern "C" int ptfnn_sync(ptfnn_sampler *s) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    return sync_and_check(s);
}

extern "C" int ptfnn_generate_draws(ptfnn_sampler *s, int32_t i0, int32_t n, float *lx, float *z, float *z_eta, float *u) {
    if (!s || n < 1 || !lx || !z || !z_eta || !u) return fail(s, PTFNN_E_INVALID, "bad argument");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const size_t rn = (size_t)s->cfg.n_replicas * n, P = s->P;
    DevBuf<float> a, b, c, d;
    CU_TRY(s, a.ensure(rn)); CU_TRY(s, b.ensure(rn * P)); CU_TRY(s, c.ensure(rn)); CU_TRY(s, d.ensure(rn));
    draws_kernel<<<dim3(n, s->cfg.n_replicas), 64, 0, s->stream>>>(s->cfg.seed, s->cfg.common_random_numbers, s->cfg.replica_offset,
                                                                  s->cfg.n_replicas, (int)P, i0, n, a.p, b.p, c.p, d.p);
    CU_TRY(s, cudaGetLastError());
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    CU_TRY(s, cudaMemcpy(lx, a.p, rn * 4, cudaMemcpyDeviceToHost)); CU_TRY(s, cudaMemcpy(z, b.p, rn * P * 4, cudaMemcpyDeviceToHost));
    CU_TRY(s, cudaMemcpy(z_eta, c.p, rn * 4, cudaMemcpyDeviceToHost)); CU_TRY(s, cudaMemcpy(u, d.p, rn * 4, cudaMemcpyDeviceToHost));
    a.release(); b.release(); c.release(); d.release();
    return PTFNN_OK;
}

extern "C" int ptfnn_swap_uniforms(const ptfnn_sampler *s, int32_t round, float *u_row) {
    if (!s || !u_row) return fail(nullptr, PTFNN_E_INVALID, "null argument");
    for (int k = 0; k + 1 < s->cfg.n_replicas_global; ++k) {
        uint32_t c[4];
        philox_draw(s->cfg.seed, (uint32_t)round, (uint32_t)k, 0u, kTagSwap, c);
        u_row[k] = u01_open_right(c[0]);
    }
    return PTFNN_OK;
}

// ------------------------------------------------------------------------------------------
// traces
// ------------------------------------------------------------------------------------------
// One batch of device -> host trace copies: every requested array is copied (2-D, rows [first, first+count) of
// every replica) into its own region of the handle's pinned staging buffer, ONE stream synchronisation,
// then widened to the reference's float64 (large blocks by a few host threads: the traces of one swap
// interval of 1024 temperatures are 16 MB).
struct TraceFetch {
    struct Item { const void *dev; size_t elem, row_elems; double *out; size_t off; bool is_int; };
    std::vector<Item> items;
    size_t bytes = 0;
    template <class T>
    void add(const T *dev, size_t row_elems, double *out, size_t R, int count) {
        if (!out) return;
        items.push_back({dev, sizeof(T), row_elems, out, bytes, std::is_integral<T>::value});
        bytes += ((R * count * row_elems * sizeof(T)) + 255) & ~(size_t)255;
    }
};

static int run_fetch(ptfnn_sampler *s, TraceFetch &f, int first, int count) {
    if (f.items.empty()) return PTFNN_OK;
    const size_t R = s->cfg.n_replicas, S = s->cfg.samples;
    if (f.bytes > s->pinned_bytes) {
        if (s->pinned) cudaFreeHost(s->pinned);
        s->pinned = nullptr; s->pinned_bytes = 0;
        CU_TRY(s, cudaHostAlloc(&s->pinned, f.bytes, cudaHostAllocDefault));
        s->pinned_bytes = f.bytes;
    }
    char *base = (char *)s->pinned;
    for (const auto &it : f.items) {
        const size_t w = count * it.row_elems * it.elem;
        CU_TRY(s, cudaMemcpy2DAsync(base + it.off, w, (const char *)it.dev + (size_t)first * it.row_elems * it.elem,
                                    S * it.row_elems * it.elem, w, R, cudaMemcpyDeviceToHost, s->stream));
    }
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    for (const auto &it : f.items) {
        const size_t n = R * count * it.row_elems;
        const char *src = base + it.off;
        double *out = it.out;
        const size_t elem = it.elem;
        const bool is_int = it.is_int;
        auto widen = [src, out, elem, is_int](size_t a, size_t b) {
            if (elem == 8) { const double *t = (const double *)src; for (size_t i = a; i < b; ++i) out[i] = t[i]; }
            else if (is_int) { const int *t = (const int *)src; for (size_t i = a; i < b; ++i) out[i] = (double)t[i]; }
            else { const float *t = (const float *)src; for (size_t i = a; i < b; ++i) out[i] = (double)t[i]; }
        };
        const size_t nthreads = n < (1u << 20) ? 1 : std::min<size_t>(std::min<unsigned>(hw, 8u), n >> 19);
        if (nthreads <= 1) { widen(0, n); continue; }
        std::vector<std::thread> th;
        const size_t chunk = (n + nthreads - 1) / nthreads;
        for (size_t k = 1; k < nthreads; ++k) th.emplace_back(widen, std::min(n, k * chunk), std::min(n, (k + 1) * chunk));
        widen(0, std::min(n, chunk));
        for (auto &t : th) t.join();
    }
    return PTFNN_OK;
}

extern "C" int ptfnn_get_traces(ptfnn_sampler *s, int32_t first, int32_t count, const ptfnn_traces *t) {
    if (!s || !t) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!s->have_state) return fail(s, PTFNN_E_STATE, "ptfnn_init_chains must come first");
    if (first < 0 || count < 1 || first + count > s->cfg.samples) return fail(s, PTFNN_E_INVALID, "rows [%d,%d) outside [0,%d)", first, first + count, s->cfg.samples);
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    int rc = sync_and_check(s);
    if (rc) return rc;
    if ((t->prior_prop || t->diff_prop || t->mh_prob || t->accepted) && !s->cfg.debug_traces)
        return fail(s, PTFNN_E_STATE, "debug traces were not enabled in the config");
    const size_t R = s->cfg.n_replicas;
    TraceFetch f;
    f.add(s->pos_w.p, (size_t)s->P, t->pos_w, R, count);
    f.add(s->lik_prop.p, 1, t->lik_prop, R, count);
    f.add(s->rmse_tr.p, 1, t->rmse_train, R, count);
    f.add(s->rmse_te.p, 1, t->rmse_test, R, count);
    f.add(s->acc_tr.p, 1, t->acc_train, R, count);
    f.add(s->acc_te.p, 1, t->acc_test, R, count);
    f.add(s->accept_list.p, 1, t->accept_list, R, count);
    if (s->cfg.debug_traces) {
        f.add(s->dbg_prior.p, 1, t->prior_prop, R, count);
        f.add(s->dbg_diff.p, 1, t->diff_prop, R, count);
        f.add(s->dbg_mh.p, 1, t->mh_prob, R, count);
    }
    if ((rc = run_fetch(s, f, first, count))) return rc;
    if (t->accepted)
        CU_TRY(s, cudaMemcpy2D(t->accepted, count, s->dbg_acc.p + first, s->cfg.samples, count, s->cfg.n_replicas, cudaMemcpyDeviceToHost));
    return PTFNN_OK;
}

extern "C" int ptfnn_traces_begin(ptfnn_sampler *s, int32_t first, int32_t count, int32_t with_pos_w, int32_t *ticket) {
    if (!s || !ticket) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!s->have_state) return fail(s, PTFNN_E_STATE, "ptfnn_init_chains must come first");
    if (first < 0 || count < 1 || first + count > s->cfg.samples) return fail(s, PTFNN_E_INVALID, "rows [%d,%d) outside [0,%d)", first, first + count, s->cfg.samples);
    if (s->device_failed) return fail(s, PTFNN_E_CUDA, "%s", kDeviceFailedMsg);
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    if (!s->copy_stream) {
        CU_TRY(s, cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        CU_TRY(s, cudaEventCreateWithFlags(&s->ev_compute, cudaEventDisableTiming));
    }
    const int k = (int)(s->fetch_seq & 1u);
    ptfnn_sampler::FetchSlot &f = s->slot[k];
    if (!f.done) CU_TRY(s, cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
    if (f.busy) CU_TRY(s, cudaEventSynchronize(f.done));          // the copy that used this slot two tickets ago
    const size_t R = s->cfg.n_replicas, S = s->cfg.samples, P = s->P;
    const size_t rc = R * (size_t)count;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    f.off[0] = take(with_pos_w ? rc * P * 4 : 0);
    for (int q = 1; q <= 5; ++q) f.off[q] = take(rc * 8);
    f.off[6] = take(rc * 4);
    if (o > f.bytes) {
        if (f.buf) cudaFreeHost(f.buf);
        f.buf = nullptr; f.bytes = 0;
        CU_TRY(s, cudaHostAlloc(&f.buf, o, cudaHostAllocDefault));
        f.bytes = o;
    }
    CU_TRY(s, cudaEventRecord(s->ev_compute, s->stream));         // everything launched so far ...
    CU_TRY(s, cudaStreamWaitEvent(s->copy_stream, s->ev_compute, 0));   // ... precedes the copies; nothing later does
    char *base = (char *)f.buf;
    auto copy = [&](size_t off, const void *dev, size_t elem, size_t row_elems) {
        const size_t w = (size_t)count * row_elems * elem;
        return cudaMemcpy2DAsync(base + off, w, (const char *)dev + (size_t)first * row_elems * elem, S * row_elems * elem, w, R,
                                 cudaMemcpyDeviceToHost, s->copy_stream);
    };
    if (with_pos_w) CU_TRY(s, copy(f.off[0], s->pos_w.p, 4, P));
    CU_TRY(s, copy(f.off[1], s->lik_prop.p, 8, 1)); CU_TRY(s, copy(f.off[2], s->rmse_tr.p, 8, 1)); CU_TRY(s, copy(f.off[3], s->rmse_te.p, 8, 1));
    CU_TRY(s, copy(f.off[4], s->acc_tr.p, 8, 1)); CU_TRY(s, copy(f.off[5], s->acc_te.p, 8, 1)); CU_TRY(s, copy(f.off[6], s->accept_list.p, 4, 1));
    CU_TRY(s, cudaEventRecord(f.done, s->copy_stream));
    f.has_w = with_pos_w != 0; f.first = first; f.count = count; f.busy = true;
    *ticket = (int32_t)s->fetch_seq;
    s->fetch_seq += 1;
    return PTFNN_OK;
}

extern "C" int ptfnn_traces_end(ptfnn_sampler *s, int32_t ticket, ptfnn_trace_views *out) {
    if (!s || !out) return fail(s, PTFNN_E_INVALID, "null argument");
    if (ticket < 0 || (unsigned int)ticket >= s->fetch_seq || (unsigned int)ticket + 2u < s->fetch_seq)
        return fail(s, PTFNN_E_STATE, "ticket %d is not one of the two most recent ptfnn_traces_begin calls", ticket);
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    ptfnn_sampler::FetchSlot &f = s->slot[ticket & 1];
    CU_TRY(s, cudaEventSynchronize(f.done));
    f.busy = false;
    const char *base = (const char *)f.buf;
    out->pos_w = f.has_w ? (const float *)(base + f.off[0]) : nullptr;
    out->lik_prop = (const double *)(base + f.off[1]); out->rmse_train = (const double *)(base + f.off[2]);
    out->rmse_test = (const double *)(base + f.off[3]); out->acc_train = (const double *)(base + f.off[4]);
    out->acc_test = (const double *)(base + f.off[5]); out->accept_list = (const int32_t *)(base + f.off[6]);
    out->first = f.first; out->count = f.count;
    return PTFNN_OK;
}

extern "C" int ptfnn_get_swap_stats(ptfnn_sampler *s, int64_t *num_swap, int64_t *total, uint8_t *swapped, int32_t max_rounds) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    { const int rc = sync_and_check(s); if (rc) return rc; }
    long long c[2] = {0, 0};
    CU_TRY(s, cudaMemcpy(c, s->swap_counters.p, 16, cudaMemcpyDeviceToHost));
    if (num_swap) *num_swap = c[0] + s->host_num_swap;
    if (total) *total = c[1] + s->host_total_prop;
    if (swapped && max_rounds > 0) {
        const size_t w = std::max(s->cfg.n_replicas_global - 1, 1);
        const size_t rounds = std::min<size_t>(max_rounds, s->swap_log.n / w);
        CU_TRY(s, cudaMemcpy(swapped, s->swap_log.p, rounds * w, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < s->host_swap_log_round.size(); ++k) {
            const size_t rd = s->host_swap_log_round[k];
            if (rd < (size_t)max_rounds) memcpy(swapped + rd * w, s->host_swap_log.data() + k * w, w);
        }
    }
    return PTFNN_OK;
}

// Result pipeline on the device traces (R:775-871 slicing, R:1036-1052 statistics).
extern "C" int ptfnn_trace_summary(ptfnn_sampler *s, int32_t first, int32_t count, ptfnn_summary *out) {
    if (!s || !out) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!s->have_state) return fail(s, PTFNN_E_STATE, "ptfnn_init_chains must come first");
    if (first < 0 || count < 1 || first + count > s->cfg.samples) return fail(s, PTFNN_E_INVALID, "rows [%d,%d) outside [0,%d)", first, first + count, s->cfg.samples);
    if ((out->w_mean == nullptr) != (out->w_std == nullptr)) return fail(s, PTFNN_E_INVALID, "w_mean and w_std go together");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    int rc = sync_and_check(s);
    if (rc) return rc;
    const int R = s->cfg.n_replicas, S = s->cfg.samples, P = s->P;
    const bool moments = out->w_mean != nullptr;
    const int cols = !moments || P <= kSumThreads ? 1 : P <= 2 * kSumThreads ? 2 : P <= 4 * kSumThreads ? 4 : 8;
    const int ctiles = moments ? (P + kSumThreads * cols - 1) / (kSumThreads * cols) : 0;
    int mblocks = std::max(1, std::min(R, 2 * s->num_sms));                   // series only: one or two short blocks per SM
    if (moments) {                                                           // persistent: two blocks per SM over all column tiles
        const int rpc = sum_rows_per_chunk(P, cols, ctiles == 1);
        const long long items = (long long)R * ((count + rpc - 1) / rpc);
        mblocks = (int)std::max<long long>(1, std::min<long long>(items, std::max(1, 2 * s->num_sms / ctiles)));
    }
    const int nblocks = moments ? ctiles * mblocks : mblocks;
    // layout of d_summary: stats[16] | ticket | acc[2P] | mean[P] | std[P] | part[nblocks*16]
    // (the ticket and the accumulators sit at offsets that do not depend on the grid: they are zero between calls)
    const size_t need = 17 + 4 * (size_t)P + 16 * (size_t)nblocks;
    if (need > s->d_summary.n) {
        CU_TRY(s, s->d_summary.ensure(need));
        CU_TRY(s, cudaMemsetAsync(s->d_summary.p, 0, need * sizeof(double), s->stream));   // acc and ticket start at zero; the kernel leaves them so
    }
    double *d_stats = s->d_summary.p, *d_ticket = d_stats + 16, *d_acc = d_ticket + 1, *d_mean = d_acc + 2 * P, *d_std = d_mean + P, *d_part = d_std + P;
    if (!s->summary_smem_opted) {                                            // per device, so per handle
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        s->summary_smem_opted = true;
    }
    cudaEvent_t e0, e1;
    CU_TRY(s, cudaEventCreate(&e0));
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return fail(s, PTFNN_E_CUDA, "cudaEventCreate"); }
    auto done = [&](int code) { cudaEventDestroy(e0); cudaEventDestroy(e1); return code; };
#define SUM_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return done(fail(s, PTFNN_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_))); } while (0)
    TraceSummaryArgs a;
    a.pos_w = s->pos_w.p;
    a.series[0] = s->rmse_tr.p; a.series[1] = s->rmse_te.p; a.series[2] = s->acc_tr.p; a.series[3] = s->acc_te.p;
    a.R = R; a.S = S; a.P = P; a.first = first; a.count = count; a.ctiles = ctiles;
    a.mblocks = mblocks; a.with_series = 1;
    a.acc = d_acc; a.part = d_part; a.ticket = reinterpret_cast<unsigned int *>(d_ticket);
    a.stats = d_stats; a.mean = d_mean; a.stdev = d_std;
    const dim3 grid(nblocks);
    const size_t smem = moments ? kSumSmemBytes : 0;
    SUM_TRY(cudaEventRecord(e0, s->stream));
    if (cols == 1) trace_summary_kernel<1><<<grid, kSumThreads, smem, s->stream>>>(a);
    else if (cols == 2) trace_summary_kernel<2><<<grid, kSumThreads, smem, s->stream>>>(a);
    else if (cols == 4) trace_summary_kernel<4><<<grid, kSumThreads, smem, s->stream>>>(a);
    else trace_summary_kernel<8><<<grid, kSumThreads, smem, s->stream>>>(a);
    SUM_TRY(cudaGetLastError());
    SUM_TRY(cudaEventRecord(e1, s->stream));
    double host[16];
    SUM_TRY(cudaMemcpyAsync(host, d_stats, sizeof host, cudaMemcpyDeviceToHost, s->stream));
    if (moments) {
        SUM_TRY(cudaMemcpyAsync(out->w_mean, d_mean, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        SUM_TRY(cudaMemcpyAsync(out->w_std, d_std, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    }
    SUM_TRY(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    SUM_TRY(cudaEventElapsedTime(&ms, e0, e1));
#undef SUM_TRY
    out->n = (int64_t)R * count;
    memcpy(out->rmse_train, host, 32); memcpy(out->rmse_test, host + 4, 32);
    memcpy(out->acc_train, host + 8, 32); memcpy(out->acc_test, host + 12, 32);
    out->kernel_ms = ms;
    out->bytes_read = (int64_t)R * count * (4 * 8 + (moments ? (int64_t)P * 4 : 0));
    return done(PTFNN_OK);
}

// Posterior-predictive moments from the device traces: forward pass of every pooled posterior
// sample (one CTA per recorded weight vector, read in place from pos_w), then the moments planes of
// trace_summary_kernel over the [samples, rows] prediction matrix.
// The [samples, rows] prediction matrix of rows [first, first+count) of every local replica's pos_w: one launch of
// the forward kernel per replica (its slice of pos_w is a batch of `count` weight vectors at stride P).
static int predictive_matrix(ptfnn_sampler *s, int32_t which, int32_t first, int32_t count, DevBuf<float> &d_fx, DevBuf<double> &d_sums,
                             long long *n_out, int *N_out) {
    if (!s->have_state) return fail(s, PTFNN_E_STATE, "ptfnn_init_chains must come first");
    if (s->cfg.task != PTFNN_TASK_REGRESSION) return fail(s, PTFNN_E_UNSUPPORTED, "predictive moments are defined for regression outputs (classification fx is a class index, C:148)");
    if (which != 0 && which != 1) return fail(s, PTFNN_E_INVALID, "which = %d (0 train, 1 test)", which);
    if (first < 0 || count < 1 || first + count > s->cfg.samples) return fail(s, PTFNN_E_INVALID, "rows [%d,%d) outside [0,%d)", first, first + count, s->cfg.samples);
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    int rc = sync_and_check(s);
    if (rc) return rc;
    const int R = s->cfg.n_replicas, S = s->cfg.samples, P = s->P;
    const int N = which == 0 ? s->n_train : s->n_test;
    const long long n = (long long)R * count;
    if (n > 0x7fffffffLL || (double)n * N * 4.0 > 16e9) return fail(s, PTFNN_E_UNSUPPORTED, "%lld samples x %d rows of predictions exceed the 16 GB staging limit: summarise a shorter slice", n, N);
    cudaError_t e;
    if ((e = d_fx.ensure((size_t)n * N + kSumTracePadFloats)) != cudaSuccess || (e = d_sums.ensure((size_t)3 * n)) != cudaSuccess)
        return fail(s, PTFNN_E_NOMEM, "cudaMalloc of the prediction matrix (%lld x %d): %s", n, N, cudaGetErrorString(e));
    const KernelSet *ks = s->ks;
    const size_t smem_fwd = (((size_t)P * 4 + 15) & ~(size_t)15) + 8 * (ks->NT / 32) * 8 + 64;
    CU_TRY(s, cudaFuncSetAttribute(ks->fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fwd));
    DataView v = which == 0 ? view(s->train_x, s->train_y, s->n_train) : view(s->test_x, s->test_y, s->n_test);
    for (int r = 0; r < R; ++r) {
        const float *wp = s->pos_w.p + ((size_t)r * S + first) * P;
        float *fxp = d_fx.p + (size_t)r * count * N, *pp = nullptr;
        double *sp = d_sums.p + (size_t)3 * r * count;
        void *args[] = {&wp, &v, &fxp, &pp, &sp};
        CU_TRY(s, cudaLaunchKernel(ks->fwd, dim3(count), dim3(ks->NT), args, smem_fwd, s->stream));
    }
    *n_out = n; *N_out = N;
    return PTFNN_OK;
}

// Percentile bands of the posterior-predictive distribution: np.percentile(fx_all, [q_lo, q_hi], axis=0)
// over the prediction matrix, on the device (quantile_bands_kernel).
extern "C" int ptfnn_predictive_bands(ptfnn_sampler *s, int32_t which, int32_t first, int32_t count, double q_lo, double q_hi,
                                      double *lo, double *hi) {
    if (!s || !lo || !hi) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!(q_lo >= 0.0 && q_lo <= q_hi && q_hi <= 100.0)) return fail(s, PTFNN_E_INVALID, "need 0 <= q_lo <= q_hi <= 100 (percent), got %g, %g", q_lo, q_hi);
    DevBuf<float> d_fx;
    DevBuf<double> d_sums, d_out;
    long long n = 0;
    int N = 0;
    int rc = predictive_matrix(s, which, first, count, d_fx, d_sums, &n, &N);
    if (rc) return rc;
    CU_TRY(s, d_out.ensure((size_t)2 * N));
    auto split = [&](double q, long long *k, double *f) {          // np.percentile, method 'linear': virtual index q/100 (n-1)
        const double idx = q / 100.0 * (double)(n - 1);
        long long kk = (long long)std::floor(idx);
        if (kk > n - 1) kk = n - 1;
        *k = kk; *f = kk >= n - 1 ? 0.0 : idx - (double)kk;
    };
    long long k0, k1;
    double f0, f1;
    split(q_lo, &k0, &f0); split(q_hi, &k1, &f1);
    quantile_bands_kernel<<<(N + kQCols - 1) / kQCols, kQThreads, 0, s->stream>>>(d_fx.p, n, N, k0, f0, k1, f1, d_out.p, d_out.p + N);
    CU_TRY(s, cudaGetLastError());
    CU_TRY(s, cudaMemcpyAsync(lo, d_out.p, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaMemcpyAsync(hi, d_out.p + N, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return PTFNN_OK;
}

extern "C" int ptfnn_predictive_summary(ptfnn_sampler *s, int32_t which, int32_t first, int32_t count,
                                        double *mean, double *stdev, double *rmse_of_mean) {
    if (!s || !mean || !stdev) return fail(s, PTFNN_E_INVALID, "null argument");
    DevBuf<float> d_fx;
    DevBuf<double> d_sums, d_out;
    long long n = 0;
    int N = 0;
    int rc = predictive_matrix(s, which, first, count, d_fx, d_sums, &n, &N);
    if (rc) return rc;
    // ---- moments over the samples for every data row
    const int cols = N <= kSumThreads ? 1 : N <= 2 * kSumThreads ? 2 : N <= 4 * kSumThreads ? 4 : 8;
    const int ctiles = (N + kSumThreads * cols - 1) / (kSumThreads * cols);
    const int rpc = sum_rows_per_chunk(N, cols, ctiles == 1);
    const long long items = (n + rpc - 1) / rpc;
    const int mblocks = (int)std::max<long long>(1, std::min<long long>(items, std::max(1, 2 * s->num_sms / ctiles)));
    const size_t need = 4 * (size_t)N + 2;
    CU_TRY(s, d_out.ensure(need));
    CU_TRY(s, cudaMemsetAsync(d_out.p, 0, need * sizeof(double), s->stream));
    if (!s->summary_smem_opted) {
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        CU_TRY(s, cudaFuncSetAttribute(trace_summary_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSumSmemBytes));
        s->summary_smem_opted = true;
    }
    TraceSummaryArgs a;
    memset(&a, 0, sizeof a);
    a.pos_w = d_fx.p;
    a.R = 1; a.S = (int)n; a.P = N; a.first = 0; a.count = (int)n; a.ctiles = ctiles; a.mblocks = mblocks; a.with_series = 0;
    a.acc = d_out.p; a.mean = d_out.p + 2 * N; a.stdev = d_out.p + 3 * N;
    a.ticket = reinterpret_cast<unsigned int *>(d_out.p + 4 * N);
    a.part = nullptr; a.stats = nullptr;
    const dim3 grid(ctiles * mblocks);
    if (cols == 1) trace_summary_kernel<1><<<grid, kSumThreads, kSumSmemBytes, s->stream>>>(a);
    else if (cols == 2) trace_summary_kernel<2><<<grid, kSumThreads, kSumSmemBytes, s->stream>>>(a);
    else if (cols == 4) trace_summary_kernel<4><<<grid, kSumThreads, kSumSmemBytes, s->stream>>>(a);
    else trace_summary_kernel<8><<<grid, kSumThreads, kSumSmemBytes, s->stream>>>(a);
    CU_TRY(s, cudaGetLastError());
    CU_TRY(s, cudaMemcpyAsync(mean, a.mean, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaMemcpyAsync(stdev, a.stdev, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    std::vector<float> y(rmse_of_mean ? (size_t)N : 0);
    if (rmse_of_mean) CU_TRY(s, cudaMemcpyAsync(y.data(), which == 0 ? s->train_y.p : s->test_y.p, (size_t)N * 4, cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    if (rmse_of_mean) {
        double acc = 0.0;
        for (int r = 0; r < N; ++r) acc += (mean[r] - (double)y[r]) * (mean[r] - (double)y[r]);
        *rmse_of_mean = std::sqrt(acc / N);
    }
    return PTFNN_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU swap round
// ------------------------------------------------------------------------------------------
extern "C" int ptfnn_swap_pending(const ptfnn_sampler *s, int32_t *pending, int32_t *is_final_round) {
    if (!s) return fail(nullptr, PTFNN_E_INVALID, "null handle");
    if (pending) *pending = s->swap_pending;
    if (is_final_round) *is_final_round = s->pending_final;
    return PTFNN_OK;
}

extern "C" int ptfnn_swap_export(ptfnn_sampler *s, void *lhood_local_dev, void *rows_local_dev) {
    if (!s || !lhood_local_dev) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!s->swap_pending) return fail(s, PTFNN_E_STATE, "no swap round pending");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    if (s->device_failed) return fail(s, PTFNN_E_CUDA, "%s", kDeviceFailedMsg);
    const size_t R = s->cfg.n_replicas, P = s->P;
    const int parity = s->rounds_done & 1;
    if (s->pending_final) {
        // exit vectors [w, eta, likelihood, ...]: the lhood field is the tempered likelihood (R:442)
        CU_TRY(s, cudaMemcpyAsync(lhood_local_dev, s->lik.p, R * 8, cudaMemcpyDeviceToDevice, s->stream));
    } else {
        CU_TRY(s, cudaMemcpyAsync(lhood_local_dev, s->pub_lhood.p + (size_t)parity * s->cfg.n_replicas_global + s->cfg.replica_offset,
                                  R * 8, cudaMemcpyDeviceToDevice, s->stream));
        if (rows_local_dev)
            CU_TRY(s, cudaMemcpyAsync(rows_local_dev, s->pub_rows.p + (size_t)parity * R * (P + kRowTail), R * (P + kRowTail) * 4, cudaMemcpyDeviceToDevice, s->stream));
    }
    return PTFNN_OK;
}

static int run_sweep(ptfnn_sampler *s, cudaStream_t st, int n, const double *lhood_dev, const float *u_host, int *d_src,
                     uint8_t *d_swapped, DevBuf<float> &d_u, int32_t *src, uint8_t *swapped, int *ns_out,
                     int kind = PTFNN_SWAP_KIND_REFERENCE, const double *temperature_dev = nullptr) {
    CU_TRY(s, d_u.ensure(std::max(n - 1, 1)));
    if (n > 1) CU_TRY(s, cudaMemcpyAsync(d_u.p, u_host, (size_t)(n - 1) * 4, cudaMemcpyHostToDevice, st));
    DevBuf<int> d_ns;
    CU_TRY(s, d_ns.ensure(1));
    const size_t smem = (size_t)n * 12;
    if (smem > 48 * 1024) CU_TRY(s, cudaFuncSetAttribute((const void *)op_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    op_sweep_kernel<<<1, 128, smem, st>>>(n, lhood_dev, d_u.p, d_src, d_swapped, d_ns.p, kind, temperature_dev);
    CU_TRY(s, cudaGetLastError());
    CU_TRY(s, cudaMemcpyAsync(src, d_src, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (n > 1 && swapped) CU_TRY(s, cudaMemcpyAsync(swapped, d_swapped, (size_t)(n - 1), cudaMemcpyDeviceToHost, st));
    int ns = 0;
    CU_TRY(s, cudaMemcpyAsync(&ns, d_ns.p, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(s, cudaStreamSynchronize(st));
    d_ns.release();
    if (ns_out) *ns_out = ns;
    return PTFNN_OK;
}

extern "C" int ptfnn_swap_plan(ptfnn_sampler *s, const void *lhood_global_dev, const float *u_row, int32_t *src, uint8_t *swapped) {
    if (!s || !lhood_global_dev || !src) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!s->swap_pending) return fail(s, PTFNN_E_STATE, "no swap round pending");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const int Rg = s->cfg.n_replicas_global;
    std::vector<float> u(std::max(Rg - 1, 1));
    if (u_row) memcpy(u.data(), u_row, (size_t)(Rg - 1) * 4);
    else ptfnn_swap_uniforms(s, s->rounds_done, u.data());
    std::vector<uint8_t> sw(std::max(Rg - 1, 1));
    int ns = 0;
    int rc = run_sweep(s, s->stream, Rg, (const double *)lhood_global_dev, u.data(), s->d_src.p, s->d_swapped.p, s->d_uswap, src, sw.data(), &ns);
    if (rc) return rc;
    if (swapped) memcpy(swapped, sw.data(), (size_t)(Rg - 1));
    s->host_num_swap += ns; s->host_total_prop += Rg - 1;
    s->host_swap_log.insert(s->host_swap_log.end(), sw.begin(), sw.end());
    s->host_swap_log_round.push_back(s->rounds_done);
    s->rounds_done += 1;
    if (s->pending_final) { s->swap_pending = false; s->pending_final = false; }   // stats only: nothing to install
    return PTFNN_OK;
}

extern "C" int ptfnn_swap_apply(ptfnn_sampler *s, const int32_t *src, const void *rows_local_dev, const void *rows_in_dev) {
    if (!s || !src || !rows_local_dev) return fail(s, PTFNN_E_INVALID, "null argument");
    if (!s->swap_pending) return PTFNN_OK;   // final round: already closed by ptfnn_swap_plan
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const int Rg = s->cfg.n_replicas_global;
    CU_TRY(s, cudaMemcpyAsync(s->d_src.p, src, (size_t)Rg * 4, cudaMemcpyHostToDevice, s->stream));
    swap_apply_kernel<<<s->cfg.n_replicas, 128, 0, s->stream>>>(s->cfg.n_replicas, s->P, s->cfg.replica_offset, s->d_src.p,
                                                                (const float *)rows_local_dev, (const float *)rows_in_dev,
                                                                s->w.p, s->eta.p, s->gd_valid.p);
    CU_TRY(s, cudaGetLastError());
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    s->swap_pending = false;
    // the chain may have ended on a swap step: the coordinator's left-over round is still due (Q9)
    if (s->step == s->cfg.samples - 1 && s->rounds_done < h_total_rounds(s)) { s->swap_pending = true; s->pending_final = true; }
    return PTFNN_OK;
}

// ------------------------------------------------------------------------------------------
// multi-GPU swap round through peer memory: the ranks exchange CUDA IPC handles of their swap windows
// once; afterwards the persistent kernel completes every round on the device (chain_kernel:
// peer_exchange_lhood + chain_sweep), with no host round trip and no collective call.
// ------------------------------------------------------------------------------------------
extern "C" int ptfnn_peer_export(ptfnn_sampler *s, void *handles /* 3 x 64 bytes */) {
    if (!s || !handles) return fail(s, PTFNN_E_INVALID, "null argument");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == PTFNN_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h[3];
    CU_TRY(s, cudaIpcGetMemHandle(&h[0], s->pub_lhood.p));
    CU_TRY(s, cudaIpcGetMemHandle(&h[1], s->pub_rows.p));
    CU_TRY(s, cudaIpcGetMemHandle(&h[2], s->peer_flags.p));
    memcpy(handles, h, sizeof h);
    return PTFNN_OK;
}

extern "C" int ptfnn_peer_connect(ptfnn_sampler *s, int32_t n_ranks, int32_t rank, const void *handles /* n_ranks x 3 x 64 bytes */) {
    if (!s || !handles) return fail(s, PTFNN_E_INVALID, "null argument");
    if (n_ranks < 1 || n_ranks > kMaxPeers || rank < 0 || rank >= n_ranks) return fail(s, PTFNN_E_INVALID, "n_ranks %d (max %d), rank %d", n_ranks, kMaxPeers, rank);
    const int R = s->cfg.n_replicas, Rg = s->cfg.n_replicas_global;
    if (Rg != R * n_ranks || s->cfg.replica_offset != rank * R) return fail(s, PTFNN_E_INVALID, "peer mode needs equal contiguous blocks: Rg %d = %d ranks x %d, offset %d", Rg, n_ranks, R, s->cfg.replica_offset);
    if (s->step != 0) return fail(s, PTFNN_E_STATE, "connect the ranks before the first step");
    CU_TRY(s, cudaSetDevice(s->cfg.device));
    const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)handles;
    for (int q = 0; q < n_ranks; ++q) {
        if (q == rank) {
            s->peer_lhood[q] = s->pub_lhood.p; s->peer_rows[q] = s->pub_rows.p; s->peer_flag_ptr[q] = s->peer_flags.p;
            continue;
        }
        void **dst[3] = {&s->peer_lhood[q], &s->peer_rows[q], &s->peer_flag_ptr[q]};
        for (int k = 0; k < 3; ++k) {
            CU_TRY(s, cudaIpcOpenMemHandle(dst[k], h[q * 3 + k], cudaIpcMemLazyEnablePeerAccess));
            s->peer_opened[q][k] = true;
        }
    }
    s->n_ranks = n_ranks; s->rank = rank;
    return PTFNN_OK;
}

// ------------------------------------------------------------------------------------------
// single operations
// ------------------------------------------------------------------------------------------
struct OpData {
    DevBuf<float> x, y, w;
    int rows = 0;
    void release() { x.release(); y.release(); w.release(); }
};

static int op_prepare(int device, int task, int I, int H, int O, const double *data, int rows, int n_cols,
                      const double *w, const KernelSet **ks, OpData &od, int batch = 1, bool labels_used = true) {
    int rc = require_device(nullptr, device);
    if (rc) return rc;
    *ks = find_kernels(task, I, H, O);
    if (!*ks) return fail(nullptr, PTFNN_E_UNSUPPORTED, "no sm_100a specialisation for task %d topology [%d,%d,%d]; built: %s", task, I, H, O, supported_list().c_str());
    if (!w) return fail(nullptr, PTFNN_E_INVALID, "null weights");
    const int P = I * H + H * O + H + O, IP = (I + 3) & ~3;
    if (data) {
        if (rows < 1 || n_cols < I + 1) return fail(nullptr, PTFNN_E_INVALID, "bad data shape [%d,%d]", rows, n_cols);
        const int bad = task == PTFNN_TASK_CLASSIFICATION ? first_bad_label(data, rows, n_cols, I, O) : -1;
        if (bad >= 0 && labels_used) return fail(nullptr, PTFNN_E_INVALID, "row %d: label %g outside [0, %d)", bad, data[(size_t)bad * n_cols + I], O);
        std::vector<float> x((size_t)rows * IP), y((size_t)((rows + 3) & ~3));
        pack_dataset(data, rows, n_cols, I, IP, x.data(), y.data());
        if (bad >= 0)                                            // evaluate_proposal never looks at the labels (C:134-153): any valid class keeps the kernel in bounds
            for (int r = 0; r < rows; ++r) y[r] = std::min(std::max(y[r], 0.0f), (float)(O - 1));
        CU_TRY(nullptr, od.x.ensure(x.size())); CU_TRY(nullptr, od.y.ensure(y.size()));
        CU_TRY(nullptr, cudaMemcpy(od.x.p, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
        CU_TRY(nullptr, cudaMemcpy(od.y.p, y.data(), y.size() * 4, cudaMemcpyHostToDevice));
        od.rows = rows;
    }
    std::vector<float> wf((size_t)P * batch);
    for (size_t j = 0; j < wf.size(); ++j) wf[j] = (float)w[j];
    CU_TRY(nullptr, od.w.ensure(wf.size()));
    CU_TRY(nullptr, cudaMemcpy(od.w.p, wf.data(), wf.size() * 4, cudaMemcpyHostToDevice));
    return PTFNN_OK;
}

// data set (padded row-major x[rows][IP]) -> UMMA A tiles of the tcgen05 likelihood pass (K5)
static int pack_a_tiles(ptfnn_sampler *s, const KernelSet *ks, const float *x, int rows, int IP, DevBuf<float> &tiles, cudaStream_t st) {
    const size_t ntiles = ((size_t)rows + 127) / 128;
    CU_TRY(s, tiles.ensure(ntiles * (size_t)ks->a_tile_floats));
    int n = rows, ip = IP; float *tp = tiles.p;
    void *args[] = {&x, &n, &ip, &tp};
    const unsigned blocks = (unsigned)std::min<size_t>(1024, (ntiles * 128 * 16 + 255) / 256);
    CU_TRY(s, cudaLaunchKernel(ks->pack_a, dim3(blocks), dim3(256), args, 0, st));
    CU_TRY(s, cudaGetLastError());
    return PTFNN_OK;
}

static int op_forward(int device, int task, int I, int H, int O, const double *data, int rows, int n_cols,
                      const double *w, double *fx, double *prob, double *sums /* [batch][3] */, int batch = 1) {
    const KernelSet *ks;
    OpData od;
    int rc = op_prepare(device, task, I, H, O, data, rows, n_cols, w, &ks, od, batch, /*labels_used=*/sums != nullptr);
    if (rc) { od.release(); return rc; }
    const int P = I * H + H * O + H + O, IP = (I + 3) & ~3;
    DevBuf<float> d_fx, d_prob;
    DevBuf<double> d_sums;
    CU_TRY(nullptr, d_fx.ensure((size_t)rows * batch)); CU_TRY(nullptr, d_sums.ensure((size_t)3 * batch));
    if (prob) CU_TRY(nullptr, d_prob.ensure((size_t)rows * O * batch));
    DataView v{od.x.p, od.y.p, rows};
    const float *wp = od.w.p; float *fxp = d_fx.p; float *pp = prob ? d_prob.p : nullptr; double *sp = d_sums.p;
    DevBuf<float> d_tiles;
    if (ks->fwd_tc) {
        // K5: wide-hidden nets go through the tcgen05 path (data set -> UMMA A tiles -> tensor-core forward)
        int rc2 = pack_a_tiles(nullptr, ks, od.x.p, rows, IP, d_tiles, 0);
        if (rc2) { od.release(); d_fx.release(); d_prob.release(); d_sums.release(); d_tiles.release(); return rc2; }
        const float *tp = d_tiles.p, *yp = od.y.p; int n = rows;
        void *args[] = {&wp, &tp, &yp, &n, &fxp, &pp, &sp};
        CU_TRY(nullptr, cudaFuncSetAttribute(ks->fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, ks->tc_smem_bytes));
        CU_TRY(nullptr, cudaLaunchKernel(ks->fwd_tc, dim3(batch), dim3(ks->NT), args, (size_t)ks->tc_smem_bytes, 0));
    } else {
        void *args[] = {&wp, &v, &fxp, &pp, &sp};
        const size_t smem = (((size_t)P * 4 + 15) & ~(size_t)15) + 8 * (ks->NT / 32) * 8 + 64;
        CU_TRY(nullptr, cudaFuncSetAttribute(ks->fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(nullptr, cudaLaunchKernel(ks->fwd, dim3(batch), dim3(ks->NT), args, smem, 0));
    }
    CU_TRY(nullptr, cudaDeviceSynchronize());
    d_tiles.release();
    if (fx) {
        std::vector<float> t((size_t)rows * batch);
        CU_TRY(nullptr, cudaMemcpy(t.data(), d_fx.p, t.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t r = 0; r < t.size(); ++r) fx[r] = t[r];
    }
    if (prob) {
        std::vector<float> t((size_t)rows * O * batch);
        CU_TRY(nullptr, cudaMemcpy(t.data(), d_prob.p, t.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < t.size(); ++k) prob[k] = t[k];
    }
    if (sums) CU_TRY(nullptr, cudaMemcpy(sums, d_sums.p, (size_t)24 * batch, cudaMemcpyDeviceToHost));
    od.release(); d_fx.release(); d_prob.release(); d_sums.release();
    return PTFNN_OK;
}

extern "C" int ptfnn_op_evaluate_proposal(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *data,
                                          int32_t rows, int32_t n_cols, const double *w, double *fx, double *prob) {
    if (!data || !fx) return fail(nullptr, PTFNN_E_INVALID, "null argument");
    return op_forward(device, task, I, H, O, data, rows, n_cols, w, fx, task == kTaskCls ? prob : nullptr, nullptr);
}

// Posterior-predictive outputs: the reference allocates fx_train_all / fx_test_all
// [samples, rows] but returns zeros (R:785-788, R:809-815 are commented out "for memory").  Here the forward
// pass of every posterior sample runs as one batch (one CTA per weight vector; tcgen05 for wide nets).
extern "C" int ptfnn_op_posterior_predictive(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *data,
                                             int32_t rows, int32_t n_cols, const double *w_samples, int32_t n_samples,
                                             double *fx_all, double *sums3) {
    if (!data || !w_samples || !fx_all || n_samples < 1) return fail(nullptr, PTFNN_E_INVALID, "bad argument");
    std::vector<double> sums((size_t)3 * n_samples);
    int rc = op_forward(device, task, I, H, O, data, rows, n_cols, w_samples, fx_all, nullptr, sums.data(), n_samples);
    if (rc) return rc;
    if (sums3) memcpy(sums3, sums.data(), sums.size() * 8);
    return PTFNN_OK;
}

// The scalar epilogue of likelihood_func (R:204-205 / C:222) on the device sums.
__global__ void op_lik_epilogue_kernel(int task, int rows, const double *sums, double tau_sq, double adapt, double *out3) {
    if (task == kTaskReg) {
        out3[0] = (-0.5 * rows * log(2.0 * 3.14159265358979323846 * tau_sq) - 0.5 * sums[0] / tau_sq) / adapt;
        out3[1] = sqrt(sums[0] / rows);
        out3[2] = 0.0;
    } else {
        out3[0] = sums[0] / adapt;
        out3[1] = sqrt(sums[1] / rows);
        out3[2] = 100.0 * (sums[2] / rows);
    }
}

extern "C" int ptfnn_op_likelihood(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *data, int32_t rows,
                                   int32_t n_cols, const double *w, double tau_sq, double adapttemp, double *out3, double *fx) {
    if (!data || !out3) return fail(nullptr, PTFNN_E_INVALID, "null argument");
    double sums[3];
    int rc = op_forward(device, task, I, H, O, data, rows, n_cols, w, fx, nullptr, sums);
    if (rc) return rc;
    DevBuf<double> d;
    CU_TRY(nullptr, d.ensure(6));
    CU_TRY(nullptr, cudaMemcpy(d.p, sums, 24, cudaMemcpyHostToDevice));
    op_lik_epilogue_kernel<<<1, 1>>>(task, rows, d.p, tau_sq, adapttemp, d.p + 3);
    CU_TRY(nullptr, cudaGetLastError());
    CU_TRY(nullptr, cudaMemcpy(out3, d.p + 3, 24, cudaMemcpyDeviceToHost));
    d.release();
    return PTFNN_OK;
}

static int op_langevin_impl(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *data,
                            int32_t rows, int32_t n_cols, const double *w, double learn_rate, int32_t depth,
                            double *w_out, int repeats, double *kernel_ms);

extern "C" int ptfnn_op_langevin_gradient(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *data,
                                          int32_t rows, int32_t n_cols, const double *w, double learn_rate, int32_t depth,
                                          double *w_out) {
    return op_langevin_impl(device, task, I, H, O, data, rows, n_cols, w, learn_rate, depth, w_out, 1, nullptr);
}

extern "C" int ptfnn_time_langevin_gradient(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *data,
                                            int32_t rows, int32_t n_cols, const double *w, double learn_rate, int32_t depth,
                                            int32_t repeats, double *kernel_ms) {
    if (!kernel_ms || repeats < 1) return fail(nullptr, PTFNN_E_INVALID, "bad argument");
    std::vector<double> out((size_t)I * H + (size_t)H * O + H + O);
    return op_langevin_impl(device, task, I, H, O, data, rows, n_cols, w, learn_rate, depth, out.data(), repeats, kernel_ms);
}

static int op_langevin_impl(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *data,
                            int32_t rows, int32_t n_cols, const double *w, double learn_rate, int32_t depth,
                            double *w_out, int repeats, double *kernel_ms) {
    if (!data || !w_out || depth < 0) return fail(nullptr, PTFNN_E_INVALID, "bad argument");
    const KernelSet *ks;
    OpData od;
    int rc = op_prepare(device, task, I, H, O, data, rows, n_cols, w, &ks, od);
    if (rc) { od.release(); return rc; }
    const int P = I * H + H * O + H + O, IP = (I + 3) & ~3;
    DevBuf<float> d_out;
    CU_TRY(nullptr, d_out.ensure(P));
    DataView v{od.x.p, od.y.p, rows};
    const float *wp = od.w.p; float *op = d_out.p; float lr = (float)learn_rate; int dep = depth;
    void *args[] = {&wp, &op, &v, &lr, &dep};
    const size_t smem = (((size_t)P * 4 + 15) & ~(size_t)15) + 16 + (size_t)2 * (kTileRows * IP * 4 + kTileRows * 4) + (size_t)team_smem_floats(H, O) * 4 + 16;
    CU_TRY(nullptr, cudaFuncSetAttribute(ks->sgd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CU_TRY(nullptr, cudaEventCreate(&e0)); CU_TRY(nullptr, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < repeats; ++rep) {
        CU_TRY(nullptr, cudaEventRecord(e0, 0));
        CU_TRY(nullptr, cudaLaunchKernel(ks->sgd, dim3(1), dim3(ks->sgd_threads), args, smem, 0));
        CU_TRY(nullptr, cudaEventRecord(e1, 0));
        CU_TRY(nullptr, cudaDeviceSynchronize());
        float ms = 0.f;
        CU_TRY(nullptr, cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = best;
    std::vector<float> t(P);
    CU_TRY(nullptr, cudaMemcpy(t.data(), d_out.p, (size_t)P * 4, cudaMemcpyDeviceToHost));
    for (int j = 0; j < P; ++j) w_out[j] = t[j];
    od.release(); d_out.release();
    return PTFNN_OK;
}

extern "C" int ptfnn_op_prior(int32_t device, int32_t task, int32_t I, int32_t H, int32_t O, const double *w, double sigma_squared,
                              double nu_1, double nu_2, double tausq, double *out) {
    if (!w || !out) return fail(nullptr, PTFNN_E_INVALID, "null argument");
    int rc = require_device(nullptr, device);
    if (rc) return rc;
    const int P = I * H + H * O + H + O;
    std::vector<float> wf(P);
    for (int j = 0; j < P; ++j) wf[j] = (float)w[j];
    DevBuf<float> dw; DevBuf<double> d;
    CU_TRY(nullptr, dw.ensure(P)); CU_TRY(nullptr, d.ensure(1));
    CU_TRY(nullptr, cudaMemcpy(dw.p, wf.data(), (size_t)P * 4, cudaMemcpyHostToDevice));
    op_prior_kernel<<<1, 128>>>(task, I, H, O, dw.p, sigma_squared, nu_1, nu_2, tausq, d.p);
    CU_TRY(nullptr, cudaGetLastError());
    CU_TRY(nullptr, cudaMemcpy(out, d.p, 8, cudaMemcpyDeviceToHost));
    dw.release(); d.release();
    return PTFNN_OK;
}

extern "C" int ptfnn_op_forward_pass(int32_t device, int32_t I, int32_t H, int32_t O, const double *x, const double *w,
                                     double *hidout, double *out) {
    if (!x || !w || !hidout || !out || I < 1 || H < 1 || O < 1) return fail(nullptr, PTFNN_E_INVALID, "bad argument");
    int rc = require_device(nullptr, device);
    if (rc) return rc;
    const int P = I * H + H * O + H + O;
    std::vector<float> buf((size_t)I + P);
    for (int i = 0; i < I; ++i) buf[i] = (float)x[i];
    for (int j = 0; j < P; ++j) buf[I + j] = (float)w[j];
    DevBuf<float> d;
    CU_TRY(nullptr, d.ensure((size_t)I + P + H + O));
    CU_TRY(nullptr, cudaMemcpy(d.p, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice));
    if ((size_t)H * 4 > 48 * 1024) CU_TRY(nullptr, cudaFuncSetAttribute((const void *)op_forward_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H * 4));
    op_forward_row_kernel<<<1, 128, (size_t)H * 4>>>(I, H, O, d.p, d.p + I, d.p + I + P, d.p + I + P + H);
    CU_TRY(nullptr, cudaGetLastError());
    std::vector<float> r((size_t)H + O);
    CU_TRY(nullptr, cudaMemcpy(r.data(), d.p + I + P, r.size() * 4, cudaMemcpyDeviceToHost));
    for (int h = 0; h < H; ++h) hidout[h] = r[h];
    for (int o = 0; o < O; ++o) out[o] = r[H + o];
    d.release();
    return PTFNN_OK;
}

extern "C" int ptfnn_op_swap_sweep_kind(int32_t device, int32_t n, const double *lhood, const float *u_row, int32_t swap_kind,
                                        const double *temperature, int32_t *src, uint8_t *swapped) {
    if (n < 1 || !lhood || !src || (n > 1 && !u_row)) return fail(nullptr, PTFNN_E_INVALID, "bad argument");
    if (swap_kind != PTFNN_SWAP_KIND_REFERENCE && swap_kind != PTFNN_SWAP_KIND_RATIO_TEMPERATURE) return fail(nullptr, PTFNN_E_INVALID, "bad swap_kind %d", swap_kind);
    if (swap_kind != PTFNN_SWAP_KIND_REFERENCE && !temperature) return fail(nullptr, PTFNN_E_INVALID, "swap_kind %d needs the temperatures", swap_kind);
    int rc = require_device(nullptr, device);
    if (rc) return rc;
    DevBuf<double> dl, dt; DevBuf<int> ds; DevBuf<uint8_t> dsw; DevBuf<float> du;
    CU_TRY(nullptr, dl.ensure(n)); CU_TRY(nullptr, ds.ensure(n)); CU_TRY(nullptr, dsw.ensure(std::max(n - 1, 1)));
    CU_TRY(nullptr, cudaMemcpy(dl.p, lhood, (size_t)n * 8, cudaMemcpyHostToDevice));
    if (temperature) {
        CU_TRY(nullptr, dt.ensure(n));
        CU_TRY(nullptr, cudaMemcpy(dt.p, temperature, (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    rc = run_sweep(nullptr, 0, n, dl.p, u_row, ds.p, dsw.p, du, src, swapped, nullptr, swap_kind, temperature ? dt.p : nullptr);
    dl.release(); dt.release(); ds.release(); dsw.release(); du.release();
    return rc;
}

extern "C" int ptfnn_op_swap_sweep(int32_t device, int32_t n, const double *lhood, const float *u_row, int32_t *src, uint8_t *swapped) {
    return ptfnn_op_swap_sweep_kind(device, n, lhood, u_row, PTFNN_SWAP_KIND_REFERENCE, nullptr, src, swapped);
}
