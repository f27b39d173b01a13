/* ptfnn.h -- C ABI of libptfnn.so: the B200 (sm_100a) parallel-tempering Bayesian-FNN sampler.
 *
 * The reference (sydney-machine-learning/parallel-tempering-neural-net) is pure Python and has NO
 * FFI of its own; its boundary is the class surface Network / ptReplica / ParallelTempering
 * (SURVEY.md section 8b).  This header is the interface a reference maintainer would bind with
 * ctypes underneath those classes (INTEGRATION.md shows the stub).  Every entry point cites the
 * reference code it replaces:
 *
 *   R: = multicore-pt-regression/pt_timeseries_regression.py
 *   C: = multicore-pt-classification/pt_classification.py
 *
 * Conventions
 *   - plain C, no C++/torch types; every call returns int (0 = PTFNN_OK, < 0 = error) and never
 *     throws across the boundary; ptfnn_last_error() gives the text.
 *   - host arrays are row-major float64, exactly the ndarrays the reference passes around
 *     (caller-owned, copied on entry); traces come back in caller-allocated float64 buffers.
 *     Device arithmetic is float32 with float64 scalar reductions ("numerics").
 *   - weight vector layout (R:80-97): [W1 (I x H row-major), W2 (H x O row-major), B1 (H), B2 (O)].
 *   - a handle is not re-entrant: one host thread per handle.  Kernels run asynchronously on the
 *     handle's stream; ptfnn_get_* / ptfnn_sync synchronise.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with PTFNN_E_CUDA.
 */
#ifndef PTFNN_H_
#define PTFNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTFNN_ABI_VERSION 2

#define PTFNN_OK 0
#define PTFNN_E_INVALID (-1)     /* bad argument / shape */
#define PTFNN_E_CUDA (-2)        /* CUDA runtime error, or no device */
#define PTFNN_E_STATE (-3)       /* call order violated (e.g. run before set_data) */
#define PTFNN_E_UNSUPPORTED (-4) /* e.g. regression with n_out != 1 (the reference requires it, R:132) */
#define PTFNN_E_NOMEM (-5)

#define PTFNN_TASK_REGRESSION 0     /* R: Gaussian likelihood, tau^2 (eta) proposals */
#define PTFNN_TASK_CLASSIFICATION 1 /* C: softmax-of-sigmoid multinomial likelihood */

#define PTFNN_SWAP_RULE_AUTO (-1)
#define PTFNN_SWAP_RULE_AFTER_I 0   /* R:427  swap when i % s == 0 and i != 0 */
#define PTFNN_SWAP_RULE_BEFORE_I1 1 /* C:438  swap when (i+1) % s == 0 */

/* The swap PROBABILITY.  The reference's scripts use R:674 / C:683; its drafts carry a temperature-aware
 * ratio rule (Misc/ldpt_fnn_multi_fixed.py:520) that no published result uses -- opt-in, single GPU, and
 * outside replay parity. */
#define PTFNN_SWAP_KIND_REFERENCE 0          /* min(1, 0.5 exp(min(709, lhood2 - lhood1))) */
#define PTFNN_SWAP_KIND_RATIO_TEMPERATURE 1  /* (lhood1 / (lhood2 or 1)) * (1/T1 * 1/T2) */

typedef struct ptfnn_sampler ptfnn_sampler;

/* Mirrors the arguments of ParallelTempering.__init__ (R:489 / C:499) and the constants hard-coded
 * in ptReplica.run (R:258-275, R:301). */
typedef struct ptfnn_config {
    int32_t abi_version;            /* PTFNN_ABI_VERSION */
    int32_t task;                   /* PTFNN_TASK_* */
    int32_t n_in, n_hidden, n_out;  /* topology [I, H, O] */
    int32_t n_replicas;             /* temperatures held by THIS handle (one handle per GPU) */
    int32_t n_replicas_global;      /* whole ladder; == n_replicas on a single GPU */
    int32_t replica_offset;         /* ladder index of this handle's first temperature */
    int32_t samples;                /* S = int(NumSample / num_chains), R:506 */
    int32_t swap_interval;          /* R:496 */
    int32_t swap_rule;              /* PTFNN_SWAP_RULE_*; AUTO picks the task's own rule */
    int32_t use_langevin_gradients; /* R:168 */
    int32_t common_random_numbers;  /* free-running mode: 1 = every replica sees the same lx / proposal
                                       noise / eta noise, as the forked reference does */
    int32_t memoize_gradient;       /* 1 = reuse langevin_gradient(w) while w is unchanged
                                       (bit-identical results) */
    int32_t device;                 /* CUDA ordinal */
    int32_t barrier_timeout_ms;     /* limit on time WITHOUT PROGRESS in a device-side wait (swap-round grid
                                     * barrier, peer flags); 0 = 20 000.  A wait that gives up ends the launch,
                                     * leaves state and traces untouched and makes every later call on the
                                     * handle fail with PTFNN_E_CUDA until ptfnn_init_chains */
    int32_t debug_traces;           /* 1 = also record prior_prop / diff_prop / mh_prob / accepted */
    int32_t speculation;            /* ladders that leave CTA slots free: CTAs per temperature that evaluate the
                                     * next steps speculatively, each as if the steps before it were rejected
                                     * (results are those of the sequential chain, bit for bit);
                                     * 0 = automatic (follows the acceptance rate and the launch times observed on
                                     * the run), 1 = off, K = that many (clamped to what is co-resident, <= 16) */
    int32_t swap_kind;              /* PTFNN_SWAP_KIND_* */
    int32_t window_plan;            /* shape of the speculative windows (measurement only; the results do not depend
                                     * on it): 0 = chosen per window (default), 1 = random-walk steps on CTAs of their
                                     * own ("apart"), 2 = before the CTA's Langevin step ("riding") */
    uint64_t seed;                  /* Philox key (free-running mode) */
    double l_prob;                  /* langevin_prob, R:174 (C:192 fixes 0.5) */
    double learn_rate;              /* R:172 */
    double step_w;                  /* 0.025, R:258 */
    double step_eta;                /* 0.2,   R:260 */
    double sigma_squared;           /* 25,    R:273 */
    double nu_1, nu_2;              /* 0, 0,  R:274-275 */
    double pt_fraction;             /* 0.6,   R:301 */
} ptfnn_config;

/* Draws for replay mode, covering steps [i0, i0+n) of every local replica (i0 = current step).
 * float32 on purpose: these are the numbers the device consumes.  SURVEY Q5 gives the order in
 * which the reference draws them. */
typedef struct ptfnn_draws {
    const float *lx;     /* [n_replicas, n]      R:327 np.random.uniform                           */
    const float *z;      /* [n_replicas, n, P]   R:331 / R:353 standard normals (proposal = loc + step_w*z) */
    const float *z_eta;  /* [n_replicas, n]      R:355 (regression only; may be NULL otherwise)    */
    const float *u;      /* [n_replicas, n]      R:387 random.uniform (MH accept)                  */
    const float *u_swap; /* [rounds_in_span, n_replicas_global-1]  R:677 coordinator uniforms      */
    int32_t n;           /* steps covered */
    int32_t n_swap_rounds; /* rows of u_swap */
} ptfnn_draws;

/* Caller-allocated float64 outputs, rows [first, first+count) of each replica's trace; any pointer
 * may be NULL.  Semantics are the reference's arrays: row 0 is the initial row
 * (pos_w = 1, likelihood = -100, rest 0), row i+1 is written by step i. */
typedef struct ptfnn_traces {
    double *pos_w;       /* [n_replicas, count, P]  R:240, R:408, R:417 */
    double *lik_prop;    /* [n_replicas, count]     likeh_list[:,0]  R:391 (tempered) / C:404 (x adapttemp) */
    double *rmse_train;  /* [n_replicas, count]     R:412 / R:420 */
    double *rmse_test;   /* [n_replicas, count] */
    double *acc_train;   /* [n_replicas, count]     C:414 (only on accept), R:403 (always 0) */
    double *acc_test;    /* [n_replicas, count] */
    double *accept_list; /* [n_replicas, count]     R:380 (# accepted BEFORE the step) */
    /* debug_traces only (not reference outputs; used by the parity tests) */
    double *prior_prop;  /* [n_replicas, count] */
    double *diff_prop;   /* [n_replicas, count] */
    double *mh_prob;     /* [n_replicas, count] */
    uint8_t *accepted;   /* [n_replicas, count] */
} ptfnn_traces;

/* ---- library ---- */
int ptfnn_abi_version(void);
const char *ptfnn_build_info(void);            /* arch, compiler, specialised topologies */
int ptfnn_device_count(void);                  /* < 0 on CUDA error, 0 if no device */
void ptfnn_default_config(ptfnn_config *cfg);  /* reference defaults (R:258-275, R:301) */
const char *ptfnn_last_error(const ptfnn_sampler *s); /* s == NULL: last error of a failed create / op */

/* ---- topologies.  The kernels are compiled specialisations of [n_in, n_hidden, n_out] (the reference
 * fixes the topology per run: R:915, C:920-986).  ptfnn_has_topology tells whether one is loaded;
 * ptfnn_register_kernels adds one that was compiled on demand from csrc/topo_inst.cu into its own
 * shared library (the Python binding does this automatically: capi.ensure_topology). */
int ptfnn_has_topology(int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out);
int ptfnn_register_kernels(const void *kernel_set /* that library's ptfnn_topology_kernels() */,
                           int32_t registry_version /* its ptfnn_topology_registry_version() */);

/* ---- lifetime: replaces ParallelTempering.__init__ + initialize_chains (R:489-527, R:639-650) ---- */
int ptfnn_create(const ptfnn_config *cfg, const double *temperatures /* [n_replicas], R:615-636 */,
                 ptfnn_sampler **out);
int ptfnn_destroy(ptfnn_sampler *s);
int ptfnn_set_stream(ptfnn_sampler *s, void *cuda_stream); /* cudaStream_t, e.g. torch's current stream */

/* traindata / testdata as passed to ParallelTempering (R:491-492): row-major [rows, n_cols] float64,
 * inputs in columns [0, I), target (R:201) or integer class label (C:210) in column I.
 * The arrays are read before the call returns (caller keeps them); the upload itself is queued on the handle's
 * stream from page-locked staging and does not wait for the device. */
int ptfnn_set_data(ptfnn_sampler *s, const double *train, int32_t n_train, const double *test,
                   int32_t n_test, int32_t n_cols);

/* w: [n_replicas, P] initial weights (R:649 np.random.randn).  Runs the pre-loop part of
 * ptReplica.run (R:266-285 / C:271-284): eta = log var(fx_train - y) (regression), tau, prior,
 * current likelihood / T, and resets traces, counters and the step index to 0. */
int ptfnn_init_chains(ptfnn_sampler *s, const double *w);

/* Overwrite / read the chain state between runs (teacher-forced replay, checkpointing).
 * Any pointer may be NULL.  lik is the current TEMPERED log-likelihood (R:397), tau the last
 * PROPOSED tau^2 (R:356). */
int ptfnn_set_state(ptfnn_sampler *s, const double *w, const double *eta, const double *lik,
                    const double *prior, const double *tau);
int ptfnn_get_state(ptfnn_sampler *s, double *w, double *eta, double *lik, double *prior, double *tau,
                    int32_t *num_accepted);
int ptfnn_get_step(const ptfnn_sampler *s, int32_t *step, int32_t *swap_rounds_done);

/* ---- the hot path: replaces ptReplica.run's loop (R:313-437 / C:313-448) for all local replicas
 * plus, on a single GPU, the coordinator's swap rounds (R:719-752) ----
 * Advances at most n_steps steps (stops at S-1).  With n_replicas_global > n_replicas the call
 * returns early right after a step at which a swap is due (*steps_done tells), and the host
 * completes the round with ptfnn_swap_* below. */
int ptfnn_run(ptfnn_sampler *s, int32_t n_steps, int32_t *steps_done);                 /* Philox draws */
int ptfnn_replay(ptfnn_sampler *s, const ptfnn_draws *d, int32_t *steps_done);        /* recorded draws */
int ptfnn_sync(ptfnn_sampler *s);

/* Dump the Philox draws free-running mode will use for steps [i0, i0+n) (verification only). */
int ptfnn_generate_draws(ptfnn_sampler *s, int32_t i0, int32_t n, float *lx, float *z, float *z_eta,
                         float *u);
int ptfnn_swap_uniforms(const ptfnn_sampler *s, int32_t round, float *u_row /* [n_replicas_global-1] */);

int ptfnn_get_traces(ptfnn_sampler *s, int32_t first, int32_t count, const ptfnn_traces *out);
/* The same rows without stalling the sampler: ptfnn_traces_begin queues the device -> host copies of rows
 * [first, first+count) behind everything launched so far, on a copy stream of the handle's own, into one of two
 * page-locked staging slots, and returns at once; later launches on the handle's stream overlap with the copy.
 * ptfnn_traces_end waits for that copy only and hands out views INTO the slot (device dtypes: float32 weights,
 * float64 series, int32 counts -- nothing is widened or copied again).  A view stays valid until the second next
 * ptfnn_traces_begin.  This is the read-back a streaming consumer (file writer, monitor) of a long run uses. */
typedef struct ptfnn_trace_views {
    const float *pos_w;          /* [n_replicas, count, P] or NULL */
    const double *lik_prop, *rmse_train, *rmse_test, *acc_train, *acc_test;   /* [n_replicas, count] */
    const int32_t *accept_list;  /* [n_replicas, count] */
    int32_t first, count;
} ptfnn_trace_views;
int ptfnn_traces_begin(ptfnn_sampler *s, int32_t first, int32_t count, int32_t with_pos_w, int32_t *ticket);
int ptfnn_traces_end(ptfnn_sampler *s, int32_t ticket, ptfnn_trace_views *out);
/* num_swap / total_swap_proposals (R:501-502, R:769) and the per-pair decisions of every round */
int ptfnn_get_swap_stats(ptfnn_sampler *s, int64_t *num_swap, int64_t *total_swap_proposals,
                         uint8_t *swapped /* [max_rounds, n_replicas_global-1] or NULL */, int32_t max_rounds);

/* ---- result pipeline on the device traces.  Pools rows [first, first+count) of every
 * local replica -- the reference's burn-in slice, R:777, R:797-824 -- and reduces them on the device:
 * what main() prints and appends to master_result_file.txt (R:1036-1052) without copying the traces
 * to the host.  Each series is {mean, np.std (population), min, max}. */
typedef struct ptfnn_summary {
    int64_t n;             /* values pooled per series: n_replicas * count */
    double rmse_train[4];  /* R:1036-1038 */
    double rmse_test[4];   /* R:1040-1042 */
    double acc_train[4];   /* C:1130-1136 */
    double acc_test[4];
    double *w_mean;        /* [P] posterior mean of every weight over the pooled rows, or NULL */
    double *w_std;         /* [P] np.std of the same, or NULL (both or neither) */
    double kernel_ms;      /* device time of the reductions (CUDA events on the handle's stream) */
    int64_t bytes_read;    /* trace bytes those kernels read */
} ptfnn_summary;
int ptfnn_trace_summary(ptfnn_sampler *s, int32_t first, int32_t count, ptfnn_summary *out);

/* Posterior-predictive moments straight from the device traces (regression).  Every recorded
 * weight vector of rows [first, first+count) of every local replica -- the pooled posterior the reference
 * would have put into fx_train_all / fx_test_all (R:785-788, R:809-815, commented out "for memory") -- is run
 * through the network on the train (which = 0) or test (which = 1) set in one batched pass, and the
 * predictions are reduced over the samples on the device: mean[r] and std[r] (np.std) of fx[:, r] for every
 * data row r.  Neither the weights nor the [samples, rows] prediction matrix leave the GPU.
 * rmse_of_mean (optional): RMSE of the posterior-mean prediction against the targets. */
int ptfnn_predictive_summary(ptfnn_sampler *s, int32_t which, int32_t first, int32_t count,
                             double *mean /* [rows] */, double *std /* [rows] */, double *rmse_of_mean /* [1] or NULL */);
/* Percentile bands of the same posterior-predictive distribution (the 5 % - 95 % uncertainty band the
 * reference's drafts plot from fx_*_all): lo[r] = np.percentile(fx[:, r], q_lo), hi[r] = np.percentile(fx[:, r],
 * q_hi) with NumPy's default linear interpolation, for every data row r, by an exact radix select over the
 * [samples, rows] prediction matrix on the device.  q in percent, 0 <= q_lo <= q_hi <= 100. */
int ptfnn_predictive_bands(ptfnn_sampler *s, int32_t which, int32_t first, int32_t count, double q_lo, double q_hi,
                           double *lo /* [rows] */, double *hi /* [rows] */);

/* Host side of the same pipeline: byte-compatible replacements of the np.savetxt / np.loadtxt calls
 * the reference spends its result phase in (R:454-481, R:794-831).  No device involved; thread-safe
 * (one file per call).  fmt: one printf conversion, as np.savetxt's fmt ('%.18e', '%1.8f', ...).
 * ptfnn_loadtxt: out == NULL reports the shape only; otherwise capacity >= rows*cols (file size / 2 always is). */
int ptfnn_savetxt(const char *path, const double *data, int64_t rows, int64_t cols, int64_t row_stride,
                  const char *fmt);
int ptfnn_loadtxt(const char *path, double *out, int64_t capacity, int64_t *rows, int64_t *cols);

/* ---- multi-GPU round (ladder partitioned over ranks).  Device pointers are owned by
 * the caller (torch tensors), so that NCCL can move them:
 *   lhood_local  [n_replicas]        float64  swap field of each local replica (R:430 / C:439)
 *   rows_local   [n_replicas, P+2]   float32  (w, eta) of each local replica; the two words behind the P weights
 *                                             are the bit pattern of the float64 eta (R:430 moves the float64)
 * After all-gathering lhood over ranks, ptfnn_swap_plan runs the reference's sequential sweep
 * (R:741-748) with the same device code every rank and returns src[k] = ladder slot whose vector
 * ends up in slot k.  Rows whose source is remote are received into rows_in (same shape as
 * rows_local, indexed by local destination slot); ptfnn_swap_apply installs them. */
/* The same round completed on the device over peer memory (NVLink / NVSwitch), no host round trip:
 * each rank exports the CUDA IPC handles of its swap window once (ptfnn_peer_export), the handles of
 * all ranks are exchanged by the caller (e.g. torch.distributed.all_gather_object) and handed to
 * ptfnn_peer_connect.  Afterwards ptfnn_run / ptfnn_replay advance through swap rounds like the
 * single-GPU kernel does; every rank must request the same number of steps.  Replaces the reference's
 * parameter queues and events between processes (R:427-437, R:730-752).
 * ptfnn_init_chains may be called again on connected handles (the arrival flags count across runs); the
 * caller must put a barrier between the ranks before the first step of the new run, so that no rank
 * publishes into a window a slower rank is still reading. */
#define PTFNN_PEER_HANDLE_BYTES 64
int ptfnn_peer_export(ptfnn_sampler *s, void *handles /* [3][PTFNN_PEER_HANDLE_BYTES] */);
int ptfnn_peer_connect(ptfnn_sampler *s, int32_t n_ranks, int32_t rank,
                       const void *handles /* [n_ranks][3][PTFNN_PEER_HANDLE_BYTES], rank order */);

int ptfnn_swap_pending(const ptfnn_sampler *s, int32_t *pending, int32_t *is_final_round);
int ptfnn_swap_export(ptfnn_sampler *s, void *lhood_local_dev, void *rows_local_dev);
int ptfnn_swap_plan(ptfnn_sampler *s, const void *lhood_global_dev, const float *u_row /* host, NULL = Philox */,
                    int32_t *src /* host [n_replicas_global] */, uint8_t *swapped /* host [n_replicas_global-1] */);
int ptfnn_swap_apply(ptfnn_sampler *s, const int32_t *src /* host [n_replicas_global] */,
                     const void *rows_local_dev, const void *rows_in_dev);

/* ---- single operations (stateless; each replaces one reference method for drop-in use and for
 * per-function parity tests).  data: row-major [rows, n_cols] float64 host array. ---- */
/* Network.ForwardPass on one input row, any topology (R:51-55 / C:49-55): hidout [H], out [O] */
int ptfnn_op_forward_pass(int32_t device, int32_t n_in, int32_t n_hidden, int32_t n_out, const double *x,
                          const double *w, double *hidout, double *out);
/* Network.evaluate_proposal (R:120-134 -> fx; C:134-153 -> fx = argmax, prob = softmax(out)) */
int ptfnn_op_evaluate_proposal(int32_t device, int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out,
                               const double *data, int32_t rows, int32_t n_cols, const double *w,
                               double *fx /* [rows] */, double *prob /* [rows, O] or NULL */);
/* Network.langevin_gradient (R:99-118 / C:114-132): depth epochs of online SGD, rows in order */
int ptfnn_op_langevin_gradient(int32_t device, int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out,
                               const double *data, int32_t rows, int32_t n_cols, const double *w,
                               double learn_rate, int32_t depth, double *w_out /* [P] */);
/* measurement helper: best-of-`repeats` device time (CUDA events) of the langevin_gradient kernel alone */
int ptfnn_time_langevin_gradient(int32_t device, int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out,
                                 const double *data, int32_t rows, int32_t n_cols, const double *w,
                                 double learn_rate, int32_t depth, int32_t repeats, double *kernel_ms);
/* ptReplica.likelihood_func (R:200-205 / C:209-222): out = {loglik/adapttemp, rmse, accuracy} */
int ptfnn_op_likelihood(int32_t device, int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out,
                        const double *data, int32_t rows, int32_t n_cols, const double *w, double tau_sq,
                        double adapttemp, double *out3, double *fx /* [rows] or NULL */);
/* ptReplica.prior_likelihood (R:215-221 / C:224-230) */
/* Posterior-predictive forward passes in one batch: fx_all[n_samples][rows] for the weight vectors
 * w_samples[n_samples][P]; sums3[n_samples][3] = per sample {sum of squared errors | sum log prob[label],
 * sum (argmax - y)^2, #correct}.  This is what the reference allocates as fx_train_all / fx_test_all but
 * returns as zeros (R:785-788, R:809-815). */
int ptfnn_op_posterior_predictive(int32_t device, int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out,
                                  const double *data, int32_t rows, int32_t n_cols, const double *w_samples,
                                  int32_t n_samples, double *fx_all, double *sums3 /* may be NULL */);
int ptfnn_op_prior(int32_t device, int32_t task, int32_t n_in, int32_t n_hidden, int32_t n_out,
                   const double *w, double sigma_squared, double nu_1, double nu_2, double tausq, double *out);
/* ParallelTempering.swap_procedure applied as the sequential sweep of run_chains (R:659-690, R:741-748) */
int ptfnn_op_swap_sweep(int32_t device, int32_t n, const double *lhood, const float *u_row,
                        int32_t *src, uint8_t *swapped);
/* the same sweep under PTFNN_SWAP_KIND_*; temperature [n] travels with the vectors (may be NULL for kind 0) */
int ptfnn_op_swap_sweep_kind(int32_t device, int32_t n, const double *lhood, const float *u_row, int32_t swap_kind,
                             const double *temperature, int32_t *src, uint8_t *swapped);

#ifdef __cplusplus
}
#endif
#endif /* PTFNN_H_ */
