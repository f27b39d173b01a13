"""``Sampler``: one libptfnn handle = the temperatures of the ladder held by one GPU.

Thin host wrapper over include/ptfnn.h used by the reference-surface classes
(regression.ParallelTempering / classification.ParallelTempering), the benchmarks and the tests.
All compute happens in csrc/ (CUDA); this module only marshals NumPy arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import TASK_CLASSIFICATION, TASK_REGRESSION  # noqa: F401


def geometric_betas(num_chains: int, maxtemp) -> np.ndarray:
    """Inverse temperatures of the geometric ladder: everything default_beta_ladder computes for the arguments
    assign_temperatures passes reduces to logspace(0, -log10(maxtemp), num_chains) (R:607)."""
    if num_chains < 1:
        raise ValueError("Invalid number of temperatures specified.")          # R:546-547
    if maxtemp is not None and maxtemp <= 1:
        raise ValueError("``Tmax`` must be greater than 1.")                   # R:544-545
    return np.logspace(0, -np.log10(maxtemp), num_chains)


def geometric_ladder(num_chains: int, maxtemp) -> np.ndarray:
    """Temperatures of ParallelTempering.assign_temperatures (R:615-636): T_k = maxtemp^(k / (num_chains - 1))."""
    return 1.0 / geometric_betas(num_chains, maxtemp)


class Sampler:
    def __init__(self, task, topology, temperatures, samples, swap_interval, *, use_langevin_gradients=True,
                 l_prob=0.5, learn_rate=0.1, seed=0, common_random_numbers=True, memoize_gradient=True,
                 device=0, debug_traces=False, swap_rule=capi.SWAP_RULE_AUTO, n_replicas_global=None,
                 replica_offset=0, step_w=0.025, step_eta=0.2, sigma_squared=25.0, nu_1=0.0, nu_2=0.0,
                 pt_fraction=0.6, stream=None, speculation=0, swap_kind=capi.SWAP_KIND_REFERENCE, barrier_timeout_ms=0, window_plan=0):
        lib = capi.load()
        capi.ensure_topology(task, topology)      # compiles a specialisation on first use of a new topology
        c = capi.default_config()
        c.task = int(task)
        c.n_in, c.n_hidden, c.n_out = (int(x) for x in topology)
        temperatures = capi.f64(temperatures)
        c.n_replicas = temperatures.shape[0]
        c.n_replicas_global = int(n_replicas_global or c.n_replicas)
        c.replica_offset = int(replica_offset)
        c.speculation = int(speculation)      # small ladders: 0 automatic, 1 off, K CTAs per temperature
        c.window_plan = int(window_plan)      # measurement only: 0 per window, 1 'apart', 2 'riding' (same results)
        c.swap_kind = int(swap_kind)          # 0 = the reference's swap probability (R:674); 1 = the drafts' temperature-aware rule
        c.barrier_timeout_ms = int(barrier_timeout_ms)   # device-side waits give up after this long without progress (0 = 20 s)
        c.samples, c.swap_interval, c.swap_rule = int(samples), int(swap_interval), int(swap_rule)
        c.use_langevin_gradients = int(bool(use_langevin_gradients))
        c.common_random_numbers = int(bool(common_random_numbers))
        c.memoize_gradient = int(bool(memoize_gradient))
        c.device, c.debug_traces, c.seed = int(device), int(bool(debug_traces)), int(seed)
        c.l_prob, c.learn_rate, c.step_w, c.step_eta = float(l_prob), float(learn_rate), float(step_w), float(step_eta)
        c.sigma_squared, c.nu_1, c.nu_2, c.pt_fraction = float(sigma_squared), float(nu_1), float(nu_2), float(pt_fraction)
        self.cfg = c
        self.task, self.topology = int(task), tuple(int(x) for x in topology)
        self.R, self.Rg, self.S = c.n_replicas, c.n_replicas_global, c.samples
        I, H, O = self.topology
        self.P = I * H + H * O + H + O
        self.temperatures = temperatures
        self._h = C.c_void_p()
        capi.check(lib.ptfnn_create(C.byref(c), capi.ptr(temperatures), C.byref(self._h)))
        self._lib = lib
        if stream is not None:
            self.set_stream(stream)

    @classmethod
    def from_oracle_config(cls, cfg, temperatures, **kw):
        """Build from an oracle.ptfnn_numpy.PTConfig (tests / smoke only)."""
        return cls(cfg.task, cfg.topology, temperatures, cfg.samples, cfg.swap_interval,
                   use_langevin_gradients=cfg.use_langevin_gradients, l_prob=cfg.l_prob, learn_rate=cfg.learn_rate,
                   step_w=cfg.step_w, step_eta=cfg.step_eta, sigma_squared=cfg.sigma_squared, nu_1=cfg.nu_1,
                   nu_2=cfg.nu_2, pt_fraction=cfg.pt_fraction, **kw)

    # ---- lifetime ----
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ptfnn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        capi.check(rc, self._h)

    def set_stream(self, stream):
        """``stream``: a raw cudaStream_t (int) or a torch.cuda.Stream."""
        handle = getattr(stream, "cuda_stream", stream)
        self._ck(self._lib.ptfnn_set_stream(self._h, C.c_void_p(int(handle))))

    # ---- inputs ----
    def set_data(self, train, test):
        train, test = capi.f64(train), capi.f64(test)
        if train.ndim != 2 or test.ndim != 2 or train.shape[1] != test.shape[1]:
            raise ValueError("train/test must be 2-D with the same number of columns")
        self._ck(self._lib.ptfnn_set_data(self._h, capi.ptr(train), train.shape[0], capi.ptr(test), test.shape[0],
                                          train.shape[1]))
        self.n_train, self.n_test = train.shape[0], test.shape[0]

    def init_chains(self, w0):
        w0 = capi.f64(w0)
        if w0.shape != (self.R, self.P):
            raise ValueError("w0 must have shape (%d, %d)" % (self.R, self.P))
        self._ck(self._lib.ptfnn_init_chains(self._h, capi.ptr(w0)))

    def set_state(self, w=None, eta=None, lik=None, prior=None, tau=None):
        arrs = [None if a is None else capi.f64(a) for a in (w, eta, lik, prior, tau)]
        self._ck(self._lib.ptfnn_set_state(self._h, *[capi.ptr(a) for a in arrs]))

    def get_state(self):
        w, eta, lik = np.zeros((self.R, self.P)), np.zeros(self.R), np.zeros(self.R)
        prior, tau, nacc = np.zeros(self.R), np.zeros(self.R), np.zeros(self.R, dtype=np.int32)
        self._ck(self._lib.ptfnn_get_state(self._h, capi.ptr(w), capi.ptr(eta), capi.ptr(lik), capi.ptr(prior),
                                           capi.ptr(tau), capi.ptr(nacc)))
        return dict(w=w, eta=eta, lik=lik, prior=prior, tau=tau, num_accepted=nacc)

    @property
    def step(self):
        s, r = C.c_int32(), C.c_int32()
        self._ck(self._lib.ptfnn_get_step(self._h, C.byref(s), C.byref(r)))
        return s.value

    @property
    def swap_rounds_done(self):
        s, r = C.c_int32(), C.c_int32()
        self._ck(self._lib.ptfnn_get_step(self._h, C.byref(s), C.byref(r)))
        return r.value

    # ---- the hot path ----
    def run(self, n_steps=None) -> int:
        """Free-running (Philox) mode; asynchronous on the handle's stream."""
        n = self.S - 1 if n_steps is None else int(n_steps)
        done = C.c_int32()
        self._ck(self._lib.ptfnn_run(self._h, n, C.byref(done)))
        return done.value

    def replay(self, draws, n_steps=None, first_round=None) -> int:
        """Replay recorded draws (arrays indexed by absolute step / round, as in oracle Draws):
        lx[R,S-1], z[R,S-1,P], z_eta[R,S-1], u[R,S-1], u_swap[rounds,Rg-1]."""
        i0 = self.step
        n = (self.S - 1 - i0) if n_steps is None else min(int(n_steps), self.S - 1 - i0)
        if n <= 0:
            return 0
        r0 = self.swap_rounds_done if first_round is None else first_round
        lx = capi.f32(np.asarray(draws.lx)[:, i0:i0 + n])
        z = capi.f32(np.asarray(draws.z)[:, i0:i0 + n])
        u = capi.f32(np.asarray(draws.u)[:, i0:i0 + n])
        z_eta = capi.f32(np.asarray(draws.z_eta)[:, i0:i0 + n]) if draws.z_eta is not None else None
        us = capi.f32(np.asarray(draws.u_swap)[r0:]) if draws.u_swap is not None and self.Rg > 1 else None
        d = capi.Draws(capi.ptr(lx), capi.ptr(z), capi.ptr(z_eta), capi.ptr(u), capi.ptr(us), n,
                       0 if us is None else us.shape[0])
        done = C.c_int32()
        self._ck(self._lib.ptfnn_replay(self._h, C.byref(d), C.byref(done)))
        return done.value

    def sync(self):
        self._ck(self._lib.ptfnn_sync(self._h))

    def generate_draws(self, i0, n):
        lx, z = np.zeros((self.R, n), np.float32), np.zeros((self.R, n, self.P), np.float32)
        z_eta, u = np.zeros((self.R, n), np.float32), np.zeros((self.R, n), np.float32)
        self._ck(self._lib.ptfnn_generate_draws(self._h, int(i0), int(n), capi.ptr(lx), capi.ptr(z), capi.ptr(z_eta),
                                                capi.ptr(u)))
        return lx, z, z_eta, u

    def swap_uniforms(self, rnd):
        u = np.zeros(max(self.Rg - 1, 1), np.float32)
        self._ck(self._lib.ptfnn_swap_uniforms(self._h, int(rnd), capi.ptr(u)))
        return u[:self.Rg - 1]

    # ---- outputs ----
    def traces(self, first=0, count=None, pos_w=True, debug=None):
        count = self.S - first if count is None else count
        debug = bool(self.cfg.debug_traces) if debug is None else debug
        out = {k: np.empty((self.R, count)) for k in ("lik_prop", "rmse_train", "rmse_test", "acc_train", "acc_test",
                                                      "accept_list")}          # filled completely by the library
        if pos_w:
            out["pos_w"] = np.empty((self.R, count, self.P))
        if debug:
            for k in ("prior_prop", "diff_prop", "mh_prob"):
                out[k] = np.zeros((self.R, count))
            out["accepted"] = np.zeros((self.R, count), dtype=np.uint8)
        t = capi.Traces(*[capi.ptr(out.get(k)) for k in ("pos_w", "lik_prop", "rmse_train", "rmse_test", "acc_train",
                                                         "acc_test", "accept_list", "prior_prop", "diff_prop",
                                                         "mh_prob", "accepted")])
        self._ck(self._lib.ptfnn_get_traces(self._h, int(first), int(count), C.byref(t)))
        if debug:
            out["accepted"] = out["accepted"].astype(bool)
        return out

    def traces_begin(self, first, count, pos_w=True) -> int:
        """Queue the read-back of rows [first, first+count) behind everything launched so far and return at once
        (copy stream + two page-locked slots inside the library); later run() calls overlap with the copy."""
        t = C.c_int32()
        self._ck(self._lib.ptfnn_traces_begin(self._h, int(first), int(count), int(bool(pos_w)), C.byref(t)))
        return t.value

    def traces_end(self, ticket) -> dict:
        """-> NumPy VIEWS into the library's staging slot (float32 pos_w, float64 series, int32 accept_list), valid
        until the second next traces_begin(); no widening, no further copy."""
        v = capi.TraceViews()
        self._ck(self._lib.ptfnn_traces_end(self._h, int(ticket), C.byref(v)))
        shape = (self.R, v.count)
        out = {k: np.ctypeslib.as_array(getattr(v, k), shape=shape)
               for k in ("lik_prop", "rmse_train", "rmse_test", "acc_train", "acc_test", "accept_list")}
        if v.pos_w:
            out["pos_w"] = np.ctypeslib.as_array(v.pos_w, shape=(self.R, v.count, self.P))
        return out

    def trace_summary(self, first=0, count=None, posterior=True):
        """The reference's result statistics (R:1036-1044: mean / np.std / min of the pooled RMSE columns
        after the burn-in slice, R:777) reduced on the device -- the traces are not copied to the host.
        -> dict: n, rmse_train / rmse_test / acc_train / acc_test = {mean, std, min, max},
        w_mean[P], w_std[P] (posterior moments of every weight), kernel_ms, bytes_read."""
        count = self.S - first if count is None else count
        sm = capi.Summary()
        wm = ws = None
        if posterior:
            wm, ws = np.empty(self.P), np.empty(self.P)
            sm.w_mean, sm.w_std = capi.ptr(wm), capi.ptr(ws)
        self._ck(self._lib.ptfnn_trace_summary(self._h, int(first), int(count), C.byref(sm)))
        out = {"n": sm.n, "kernel_ms": sm.kernel_ms, "bytes_read": sm.bytes_read, "w_mean": wm, "w_std": ws}
        for k in ("rmse_train", "rmse_test", "acc_train", "acc_test"):
            v = getattr(sm, k)
            out[k] = {"mean": v[0], "std": v[1], "min": v[2], "max": v[3]}
        return out

    def predictive_summary(self, which="test", first=0, count=None, bands=None):
        """Posterior-predictive mean and std of every data row over the pooled posterior samples (rows
        [first, first+count) of every chain's pos_w), computed from the device traces: what np.mean / np.std
        over the fx_*_all arrays the reference comments out (R:785-788, R:809-815) would give.  Regression.
        ``bands`` = (q_lo, q_hi) in percent, e.g. (5, 95): also np.percentile(fx_all, q, axis=0) of the same
        matrix (exact radix select on the device) as ``lo`` / ``hi``.
        -> dict(mean[rows], std[rows], rmse_of_mean[, lo[rows], hi[rows]])"""
        count = self.S - first if count is None else count
        rows = self.n_train if which == "train" else self.n_test
        wh = 0 if which == "train" else 1
        mean, std, rm = np.empty(rows), np.empty(rows), C.c_double()
        self._ck(self._lib.ptfnn_predictive_summary(self._h, wh, int(first), int(count),
                                                    capi.ptr(mean), capi.ptr(std), C.byref(rm)))
        out = {"mean": mean, "std": std, "rmse_of_mean": rm.value}
        if bands is not None:
            out["lo"], out["hi"] = self.predictive_bands(which, first, count, *bands)
        return out

    def predictive_bands(self, which="test", first=0, count=None, q_lo=5.0, q_hi=95.0):
        """-> (lo[rows], hi[rows]) = np.percentile of the posterior-predictive samples of every data row."""
        count = self.S - first if count is None else count
        rows = self.n_train if which == "train" else self.n_test
        lo, hi = np.empty(rows), np.empty(rows)
        self._ck(self._lib.ptfnn_predictive_bands(self._h, 0 if which == "train" else 1, int(first), int(count),
                                                  C.c_double(q_lo), C.c_double(q_hi), capi.ptr(lo), capi.ptr(hi)))
        return lo, hi

    def swap_stats(self, max_rounds=None):
        """-> (num_swap, total_swap_proposals, swapped[rounds, Rg-1])"""
        ns, tot = C.c_int64(), C.c_int64()
        rounds = self.swap_rounds_done if max_rounds is None else max_rounds
        sw = np.zeros((max(rounds, 1), max(self.Rg - 1, 1)), dtype=np.uint8)
        self._ck(self._lib.ptfnn_get_swap_stats(self._h, C.byref(ns), C.byref(tot), capi.ptr(sw), rounds))
        return ns.value, tot.value, sw[:rounds, :self.Rg - 1].astype(bool)

    # ---- multi-GPU ladder through peer memory: rounds complete on the device ----
    def peer_export(self) -> bytes:
        """CUDA IPC handles of this rank's swap window (lhood fields, (w, eta) rows, arrival flags)."""
        buf = C.create_string_buffer(3 * capi.PEER_HANDLE_BYTES)
        self._ck(self._lib.ptfnn_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, handles, rank: int):
        """handles: the peer_export() bytes of every rank, in rank order (this rank's own entry is ignored)."""
        blob = b"".join(handles)
        if len(blob) != len(handles) * 3 * capi.PEER_HANDLE_BYTES:
            raise ValueError("each rank contributes %d bytes of IPC handles" % (3 * capi.PEER_HANDLE_BYTES))
        self._ck(self._lib.ptfnn_peer_connect(self._h, len(handles), int(rank), C.c_char_p(blob)))

    # ---- multi-GPU round completed by the host (device pointers come from torch tensors) ----
    def swap_pending(self):
        p, f = C.c_int32(), C.c_int32()
        self._ck(self._lib.ptfnn_swap_pending(self._h, C.byref(p), C.byref(f)))
        return bool(p.value), bool(f.value)

    def swap_export(self, lhood_local_ptr, rows_local_ptr):
        self._ck(self._lib.ptfnn_swap_export(self._h, C.c_void_p(lhood_local_ptr), C.c_void_p(rows_local_ptr)))

    def swap_plan(self, lhood_global_ptr, u_row=None):
        src = np.zeros(self.Rg, dtype=np.int32)
        sw = np.zeros(max(self.Rg - 1, 1), dtype=np.uint8)
        u = None if u_row is None else capi.f32(u_row)
        self._ck(self._lib.ptfnn_swap_plan(self._h, C.c_void_p(lhood_global_ptr), capi.ptr(u), capi.ptr(src),
                                           capi.ptr(sw)))
        return src, sw[:self.Rg - 1].astype(bool)

    def swap_apply(self, src, rows_local_ptr, rows_in_ptr):
        src = np.ascontiguousarray(src, dtype=np.int32)
        self._ck(self._lib.ptfnn_swap_apply(self._h, capi.ptr(src), C.c_void_p(rows_local_ptr), C.c_void_p(rows_in_ptr)))
