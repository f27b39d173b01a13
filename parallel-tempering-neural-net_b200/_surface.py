"""Shared implementation of the reference's class surface on top of libptfnn.

The two public modules ``regression`` and ``classification`` bind these bases to the signatures of
  R: = multicore-pt-regression/pt_timeseries_regression.py
  C: = multicore-pt-classification/pt_classification.py
(SURVEY 8b).  Method names, argument meaning, return tuples, output files and error behaviour are
the reference's; every number is computed by the CUDA library (capi / Sampler).  There is no
``multiprocessing`` here: the R replica processes + coordinator of the reference are one persistent
kernel launch (``ParallelTempering.run_chains``).
"""
from __future__ import annotations

import os

import numpy as np

from . import capi
from .sampler import Sampler, geometric_betas

REGRESSION, CLASSIFICATION = capi.TASK_REGRESSION, capi.TASK_CLASSIFICATION


# ==========================================================================================
# Network (R:27-134 / C:26-153)
# ==========================================================================================
class NetworkBase:
    TASK = REGRESSION

    def __init__(self, Topo, Train, Test, learn_rate):
        self.Top = Topo                      # NN topology [input, hidden, output]
        self.TrainData = Train
        self.TestData = Test
        self.lrate = learn_rate
        I, H, O = Topo
        # same initialisation calls, in the same order, as R:35-38 (keeps the caller's NumPy stream aligned)
        self.W1 = np.random.randn(I, H) / np.sqrt(I)
        self.B1 = np.random.randn(1, H) / np.sqrt(H)
        self.W2 = np.random.randn(H, O) / np.sqrt(H)
        self.B2 = np.random.randn(1, O) / np.sqrt(H)
        self.hidout = np.zeros((1, H))
        self.out = np.zeros((1, O))
        self.pred_class = 0
        self.device = 0

    def sigmoid(self, x):
        return 1 / (1 + np.exp(-x))

    def sampleEr(self, actualout):
        error = np.subtract(self.out, actualout)
        return np.sum(np.square(error)) / self.Top[2]

    # -- weight vector layout a1: [W1 (I x H), W2 (H x O), B1 (H), B2 (O)]  (R:80-97)
    def decode(self, w):
        I, H, O = self.Top
        w = np.asarray(w)
        a, b = I * H, I * H + H * O
        self.W1 = np.reshape(w[0:a], (I, H))
        self.W2 = np.reshape(w[a:b], (H, O))
        if self.TASK == CLASSIFICATION:          # C:94-95 keeps the biases 2-D
            self.B1 = w[b:b + H].reshape(1, H)
            self.B2 = w[b + H:b + H + O].reshape(1, O)
        else:
            self.B1 = w[b:b + H]
            self.B2 = w[b + H:b + H + O]

    def encode(self):
        return np.concatenate([np.ravel(self.W1), np.ravel(self.W2), np.ravel(self.B1), np.ravel(self.B2)])

    def ForwardPass(self, X):
        """R:51-55 / C:49-55 on the device (single row)."""
        hid, out = capi.op_forward_pass(self.Top, np.asarray(X, dtype=np.float64).reshape(-1), self.encode(),
                                        device=self.device)
        if self.TASK == CLASSIFICATION:
            self.hidout, self.out = hid.reshape(1, -1), out.reshape(1, -1)
            self.pred_class = int(np.argmax(self.out))
        else:
            self.hidout, self.out = hid, out

    def BackwardPass(self, Input, desired):
        """R:57-78 / C:72-82: one online-SGD step on this row == langevin_gradient over a 1-row dataset."""
        row = np.concatenate([np.asarray(Input, dtype=np.float64).reshape(-1),
                              np.asarray(desired, dtype=np.float64).reshape(-1)])[None, :]
        self.decode(capi.op_langevin_gradient(self.TASK, tuple(self.Top), row, self.encode(), self.lrate, 1,
                                              device=self.device))

    def langevin_gradient(self, data, w, depth):
        """R:99-118 / C:114-132.  Like the reference, ``w`` is updated IN PLACE (decode() makes views,
        which is why callers pass ``w.copy()``, R:330) and the updated vector is returned."""
        out = capi.op_langevin_gradient(self.TASK, tuple(self.Top), data, w, self.lrate, depth, device=self.device)
        if isinstance(w, np.ndarray) and w.dtype == np.float64 and w.flags.writeable:
            w[...] = out
        self.decode(out)
        return out

    def evaluate_proposal(self, data, w):
        """R:120-134 -> fx ;  C:134-153 -> (fx = argmax, prob = softmax of the sigmoid outputs)."""
        self.decode(np.asarray(w, dtype=np.float64))
        return capi.op_evaluate_proposal(self.TASK, tuple(self.Top), data, w, device=self.device)


# ==========================================================================================
# ptReplica (R:138-485 / C:157-494)
# ==========================================================================================
class ReplicaBase:
    TASK = REGRESSION
    NETWORK = NetworkBase

    def _init_common(self, use_langevin_gradients, learn_rate, w, minlim_param, maxlim_param, samples, traindata,
                     testdata, topology, burn_in, temperature, swap_interval, langevin_prob, path, parameter_queue,
                     main_process, event):
        self.processID = temperature
        self.parameter_queue = parameter_queue
        self.signal_main = main_process
        self.event = event
        self.temperature = temperature
        self.adapttemp = temperature
        self.swap_interval = swap_interval
        self.path = path
        self.burn_in = burn_in
        self.samples = samples
        self.topology = topology
        self.traindata = traindata
        self.testdata = testdata
        self.w = w
        self.minY = np.zeros((1, 1))
        self.maxY = np.zeros((1, 1))
        self.minlim_param = minlim_param
        self.maxlim_param = maxlim_param
        self.use_langevin_gradients = use_langevin_gradients
        self.sgd_depth = 1                       # always should be 1 (R:170)
        self.learn_rate = learn_rate
        self.l_prob = langevin_prob
        self.w_size = 0
        self.device = 0
        self.seed = None

    # Process-like shims so code that treated replicas as processes keeps working
    def start(self):
        self.run()

    def join(self, timeout=None):
        return None

    def is_alive(self):
        return False

    def rmse(self, pred, actual):
        return np.sqrt(((pred - actual) ** 2).mean())

    def accuracy(self, pred, actual):
        count = 0
        for i in range(pred.shape[0]):
            if pred[i] == actual[i]:
                count += 1
        return 100 * (count / pred.shape[0])

    def _likelihood(self, fnn, data, w, tau_sq):
        lik, rm, acc, fx = capi.op_likelihood(self.TASK, tuple(self.topology), data, w, tau_sq, self.adapttemp,
                                              device=self.device)
        return [lik, fx, rm]

    def _prior(self, sigma_squared, nu_1, nu_2, w, tausq):
        return capi.op_prior(self.TASK, tuple(self.topology), w, sigma_squared, nu_1, nu_2, tausq, device=self.device)

    def run(self):
        """The chain of THIS replica alone (no partner to swap with), on the GPU, writing the same
        per-chain files as R:454-481.  ParallelTempering.run_chains() runs all replicas at once."""
        P = _num_param(self.topology)
        self.w_size = P
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if self.seed is None else int(self.seed)
        with Sampler(self.TASK, self.topology, [self.temperature], self.samples, self.swap_interval,
                     use_langevin_gradients=self.use_langevin_gradients, l_prob=self.l_prob,
                     learn_rate=self.learn_rate, seed=seed, device=self.device) as s:
            s.set_data(self.traindata, self.testdata)
            s.init_chains(np.asarray(self.w, dtype=np.float64)[None, :])
            s.run()
            t = s.traces()
            st = s.get_state()
        _write_chain_files(self.path, self.temperature, self.TASK, self.samples, t, 0, int(st["num_accepted"][0]))
        self.w = st["w"][0]
        if self.parameter_queue is not None:       # R:442-444: final vector for the coordinator
            lik = float(st["lik"][0])
            self.parameter_queue.put(np.concatenate([self.w, [float(st["eta"][0])], [lik], [self.adapttemp],
                                                     [self.samples - 2]]))
        if self.signal_main is not None:
            self.signal_main.set()                 # R:485


def _num_param(topology):
    return topology[0] * topology[1] + topology[1] * topology[2] + topology[1] + topology[2]


def _write_chain_files(path, temperature, task, samples, t, k, num_accepted):
    """Per-chain output files, names and formats of R:454-481 / C:465-492.  capi.savetxt writes the bytes
    np.savetxt would (these calls dominate run_chains() once sampling is fast)."""
    T = str(temperature)
    savetxt = capi.savetxt
    savetxt(path + '/posterior/pos_w/' + 'chain_' + T + '.txt', t["pos_w"][k])
    fmt = '%1.8f' if task == REGRESSION else '%1.2f'                              # R:462-464 | C:473-475
    savetxt(path + '/predictions/rmse_test_chain_' + T + '.txt', t["rmse_test"][k], fmt=fmt)
    savetxt(path + '/predictions/rmse_train_chain_' + T + '.txt', t["rmse_train"][k], fmt=fmt)
    savetxt(path + '/predictions/acc_test_chain_' + T + '.txt', t["acc_test"][k], fmt='%1.2f')
    savetxt(path + '/predictions/acc_train_chain_' + T + '.txt', t["acc_train"][k], fmt='%1.2f')
    likeh = np.zeros((samples, 2))
    likeh[:, 0] = t["lik_prop"][k]
    likeh[0, :] = [-100, -100]                                                    # R:293
    savetxt(path + '/posterior/pos_likelihood/chain_' + T + '.txt', likeh, fmt='%1.4f')
    accept_ratio = num_accepted / (samples * 1.0) * 100                           # R:447
    savetxt(path + '/posterior/accept_list/chain_' + T + '_accept.txt', [accept_ratio], fmt='%1.4f')
    savetxt(path + '/posterior/accept_list/chain_' + T + '.txt', t["accept_list"][k], fmt='%1.4f')


def _per_chain(fn, n):
    """fn(k) for every chain, one thread each (the text conversion runs in the library, outside the GIL;
    the reference gets the same parallelism from its one-process-per-replica layout)."""
    if n <= 1:
        return [fn(k) for k in range(n)]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(n, os.cpu_count() or 4, 32)) as ex:
        return list(ex.map(fn, range(n)))


# ==========================================================================================
# ParallelTempering (R:487-875 / C:497-897)
# ==========================================================================================
class _Slot:
    """Stand-in for the per-replica multiprocessing.Queue of the reference (R:509)."""

    def __init__(self):
        self._items = []

    def put(self, item):
        self._items.append(item)

    def get(self):
        return self._items.pop(0)

    def empty(self):
        return not self._items


class _Flag:
    def __init__(self):
        self._set = False

    def set(self):
        self._set = True

    def clear(self):
        self._set = False

    def is_set(self):
        return self._set

    def wait(self, timeout=None):
        return self._set


class ParallelTemperingBase:
    TASK = REGRESSION
    REPLICA = ReplicaBase

    def _init_common(self, use_langevin_gradients, learn_rate, traindata, testdata, topology, num_chains, maxtemp,
                     NumSample, swap_interval, langevin_prob, path):
        self.traindata = traindata
        self.testdata = testdata
        self.topology = topology
        self.num_param = _num_param(topology)
        self.swap_interval = swap_interval
        self.path = path
        self.maxtemp = maxtemp
        self.langevin_prob = langevin_prob
        self.num_swap = 0
        self.total_swap_proposals = 0
        self.num_chains = num_chains
        self.chains = []
        self.temperatures = []
        self.NumSamples = int(NumSample / self.num_chains)                        # R:506
        self.sub_sample_size = max(1, int(0.05 * self.NumSamples))
        self.parameter_queue = [_Slot() for _ in range(num_chains)]
        self.chain_queue = _Slot()
        self.wait_chain = [_Flag() for _ in range(num_chains)]
        self.event = [_Flag() for _ in range(num_chains)]
        self.all_param = None
        self.geometric = True
        self.minlim_param = 0.0
        self.maxlim_param = 0.0
        self.minY = np.zeros((1, 1))
        self.maxY = np.ones((1, 1))
        self.model_signature = 0.0
        self.learn_rate = learn_rate
        self.use_langevin_gradients = use_langevin_gradients
        # ---- knobs that do not exist in the reference (all default to reference behaviour)
        self.device = 0
        self.seed = None                    # None: drawn from np.random, so np.random.seed() makes runs repeatable
        self.common_random_numbers = True   # forked replicas share the NumPy stream
        self.memoize_gradient = True
        self.write_files = True             # per-chain txt files of R:454-481
        self.results_from_files = True      # show_results() re-reads them (R:794-831); False = in-memory, full precision
        self.posterior_predictive = False   # True: fx_train_all / fx_test_all hold the predictions of every posterior
                                            # sample (one batched GPU pass) instead of the reference's zeros (R:785-788)
        self.predictive_moments = False     # True (regression): posterior-predictive mean / std of every train and test row over
                                            # the pooled posterior, reduced on the device, left in self.predictive
        self.predictive_bands = None        # e.g. (5, 95): with predictive_moments, also np.percentile bands of the same predictions
        self.predictive = None
        self.swap_kind = 0                  # 0: the swap probability of R:674; 1: the drafts' temperature-aware rule
                                            # (Misc/ldpt_fnn_multi_fixed.py:520) -- opt-in, outside replay parity
        self.last_sampler_seconds = None
        self.summary = None                 # device-reduced statistics of the last run (Sampler.trace_summary)
        self._traces = None

    def default_beta_ladder(self, ndim, ntemps, Tmax):
        """Inverse temperatures of the ladder (R:529-613).  For every argument list the reference's version
        survives, its result is the geometric spacing logspace(0, -log10(Tmax), ntemps) (R:607): it rejects
        bad arguments first (R:538-547), then iterates ``range(Tmax)`` over a step that divides by
        ``ntemps - 1`` (R:570-579) -- so a non-integer Tmax (None and inf included) is a TypeError and a
        one-rung ladder a ZeroDivisionError -- and the ptemcee step size it derives from ``ndim`` is never
        used once both ntemps and Tmax are given.  Same checks, same errors, same numbers here."""
        import operator
        if type(ndim) != int or ndim < 1:
            raise ValueError('Invalid number of dimensions specified.')
        if ntemps is None and Tmax is None:
            raise ValueError('Must specify one of ``ntemps`` and ``Tmax``.')
        if Tmax is not None and Tmax <= 1:
            raise ValueError('``Tmax`` must be greater than 1.')
        if ntemps is not None and (type(ntemps) != int or ntemps < 1):
            raise ValueError('Invalid number of temperatures specified.')
        operator.index(Tmax)                       # R:576 range(maxtemp): TypeError unless an integer
        if ntemps is None:
            raise TypeError("unsupported operand type(s) for ** or pow(): 'NoneType' and 'float'")   # R:577
        if ntemps == 1:
            raise ZeroDivisionError('division by zero')                                            # R:577
        return geometric_betas(ntemps, Tmax)

    def assign_temperatures(self):
        if self.geometric is True:                                                # R:624-628
            betas = self.default_beta_ladder(2, ntemps=self.num_chains, Tmax=self.maxtemp)
            for i in range(0, self.num_chains):
                self.temperatures.append(np.inf if betas[i] == 0 else 1.0 / betas[i])
        else:                                                                     # R:630-636
            tmpr_rate = (self.maxtemp / self.num_chains)
            temp = 1
            for i in range(0, self.num_chains):
                self.temperatures.append(temp)
                temp += tmpr_rate

    def _make_replica(self, w, i):
        raise NotImplementedError

    def initialize_chains(self, burn_in):
        self.burn_in = burn_in
        self.assign_temperatures()
        self.minlim_param = np.repeat([-100], self.num_param)
        self.maxlim_param = np.repeat([100], self.num_param)
        for i in range(0, self.num_chains):
            w = np.random.randn(self.num_param)                                   # R:649
            self.chains.append(self._make_replica(w, i))

    def surr_procedure(self, queue):
        if queue.empty() is False:
            return queue.get()
        return

    def swap_procedure(self, parameter_queue_1, parameter_queue_2):
        """R:659-690.  The decision min(1, 0.5*exp(min(709, l2 - l1))) is taken by the device sweep on the
        two lhood fields; the uniform comes from the caller's NumPy stream as in the reference (R:677)."""
        param1 = parameter_queue_1.get()
        param2 = parameter_queue_2.get()
        lhood1 = param1[self.num_param + 1]
        lhood2 = param2[self.num_param + 1]
        u = np.random.uniform(0, 1)
        _, sw = capi.op_swap_sweep([lhood1, lhood2], [u], device=self.device)
        swapped = bool(sw[0])
        self.total_swap_proposals += 1
        if swapped:
            self.num_swap += 1
            param1, param2 = param2, param1
        return param1, param2, swapped

    def _run_sampler(self, want_traces):
        """The sampling phase of run_chains (R:709-759): all replicas and every swap round are ONE persistent
        kernel launch.  Also leaves the device-reduced result statistics in ``self.summary``."""
        import time
        if not self.chains:
            raise RuntimeError("initialize_chains(burn_in) must be called first")
        S = self.NumSamples
        burnin = int(S * self.burn_in)
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if self.seed is None else int(self.seed)
        l_prob = self.langevin_prob if self.TASK == REGRESSION else 0.5           # C:192
        t0 = time.perf_counter()
        with Sampler(self.TASK, self.topology, self.temperatures, S, self.swap_interval,
                     use_langevin_gradients=self.use_langevin_gradients, l_prob=l_prob, learn_rate=self.learn_rate,
                     seed=seed, common_random_numbers=self.common_random_numbers,
                     memoize_gradient=self.memoize_gradient, device=self.device, swap_kind=self.swap_kind) as s:
            s.set_data(self.traindata, self.testdata)
            s.init_chains(np.stack([np.asarray(c.w, dtype=np.float64) for c in self.chains]))
            s.run()
            self.summary = s.trace_summary(burnin, S - burnin) if S > burnin else None
            if self.predictive_moments and self.TASK == REGRESSION and S > burnin:
                self.predictive = {k: s.predictive_summary(k, burnin, S - burnin, bands=self.predictive_bands)
                                   for k in ("train", "test")}
            t = s.traces() if want_traces else s.traces(first=S - 1, count=1, pos_w=False)
            st = s.get_state()
            ns, tot, _ = s.swap_stats()
        self.last_sampler_seconds = time.perf_counter() - t0
        self.num_swap += ns
        self.total_swap_proposals += tot
        for k, c in enumerate(self.chains):
            c.w = st["w"][k]
        return t, st

    def run_chains(self):
        """R:694-771."""
        open(self.path + '/num_exchange.txt', 'a').close()                        # R:704 (opened, never written)
        S = self.NumSamples
        t, st = self._run_sampler(True)
        self._traces, self._state = t, st
        if self.write_files:
            _per_chain(lambda k: _write_chain_files(self.path, self.temperatures[k], self.TASK, S, t, k,
                                                    int(st["num_accepted"][k])), self.num_chains)
        pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, likelihood_vec, accept_vec, accept = \
            self.show_results()
        print("NUMBER OF SWAPS =", self.num_swap)
        swap_perc = self.num_swap * 100 / self.total_swap_proposals              # ZeroDivisionError if no round ran, as R:769
        return (pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, likelihood_vec, swap_perc,
                accept_vec, accept)

    def run_summary(self):
        """Not in the reference: run_chains() for callers that only want what main() reports.
        The burn-in slice is pooled and reduced on the device (R:777, R:1036-1044, C:1130-1136), so neither the
        S x P traces nor any txt file leave the GPU.  -> dict with the series statistics
        ({mean, std, min, max} of rmse_train / rmse_test / acc_train / acc_test), the posterior mean / std of
        every weight, swap_perc (R:769) and accept_per (R:1009-1011)."""
        S = self.NumSamples
        t, st = self._run_sampler(False)
        out = dict(self.summary)
        out["swap_perc"] = self.num_swap * 100 / self.total_swap_proposals
        out["accept_per"] = float(np.mean(t["accept_list"][:, -1] / S) * 100)     # Q16: the count BEFORE the last step
        out["num_accepted"] = st["num_accepted"]
        return out

    # -- R:775-871 / C:780-893
    def _lik_rows(self, burnin):
        return slice(1, None) if self.TASK == REGRESSION else slice(burnin, None)  # R:801 | C:809 (Q14)

    def show_results(self):
        S, R = self.NumSamples, self.num_chains
        burnin = int(S * self.burn_in)
        nlik = S - 1 if self.TASK == REGRESSION else S - burnin
        likelihood_rep = np.zeros((R, nlik, 2))
        accept_percent = np.zeros((R, 1))
        accept_list = np.zeros((R, S))
        pos_w = np.zeros((R, S - burnin, self.num_param))
        fx_train_all = np.zeros((R, S - burnin, self.traindata.shape[0]))        # returned as zeros (R:785, R:809-815)
        rmse_train = np.zeros((R, S - burnin))
        acc_train = np.zeros((R, S - burnin))
        fx_test_all = np.zeros((R, S - burnin, self.testdata.shape[0]))
        rmse_test = np.zeros((R, S - burnin))
        acc_test = np.zeros((R, S - burnin))
        from_files = self.results_from_files and self.write_files
        t = self._traces
        if not from_files and t is None:
            raise RuntimeError("run_chains() has not produced traces yet")

        def fill(i):
            T = str(self.temperatures[i])
            if from_files:                                                        # R:794-831, read back as written
                loadtxt = capi.loadtxt
                pos_w[i, :, :] = loadtxt(self.path + '/posterior/pos_w/' + 'chain_' + T + '.txt')[burnin:, :]
                likelihood_rep[i, :] = loadtxt(self.path + '/posterior/pos_likelihood/' + 'chain_' + T + '.txt')[self._lik_rows(burnin)]
                accept_list[i, :] = loadtxt(self.path + '/posterior/accept_list/' + 'chain_' + T + '.txt')
                rmse_test[i, :] = loadtxt(self.path + '/predictions/rmse_test_chain_' + T + '.txt')[burnin:]
                rmse_train[i, :] = loadtxt(self.path + '/predictions/rmse_train_chain_' + T + '.txt')[burnin:]
                acc_test[i, :] = loadtxt(self.path + '/predictions/acc_test_chain_' + T + '.txt')[burnin:]
                acc_train[i, :] = loadtxt(self.path + '/predictions/acc_train_chain_' + T + '.txt')[burnin:]
            else:
                pos_w[i] = t["pos_w"][i, burnin:]
                likelihood_rep[i, :, 0] = t["lik_prop"][i, self._lik_rows(burnin)]
                accept_list[i] = t["accept_list"][i]
                rmse_test[i], rmse_train[i] = t["rmse_test"][i, burnin:], t["rmse_train"][i, burnin:]
                acc_test[i], acc_train[i] = t["acc_test"][i, burnin:], t["acc_train"][i, burnin:]

        _per_chain(fill, R)
        if self.posterior_predictive:
            for i in range(R):
                fx_train_all[i] = capi.op_posterior_predictive(self.TASK, self.topology, self.traindata, pos_w[i], self.device)[0]
                fx_test_all[i] = capi.op_posterior_predictive(self.TASK, self.topology, self.testdata, pos_w[i], self.device)[0]
        posterior = pos_w.transpose(2, 0, 1).reshape(self.num_param, -1)
        likelihood_vec = likelihood_rep.transpose(2, 0, 1).reshape(2, -1)
        rmse_train = rmse_train.reshape(R * (S - burnin), 1)
        acc_train = acc_train.reshape(R * (S - burnin), 1)
        rmse_test = rmse_test.reshape(R * (S - burnin), 1)
        acc_test = acc_test.reshape(R * (S - burnin), 1)
        accept_vec = accept_list
        accept = np.sum(accept_percent) / R                                       # always 0: never filled (R:780, R:860)
        capi.savetxt(self.path + '/likelihood.txt', likelihood_vec.T, fmt='%1.5f')
        capi.savetxt(self.path + '/accept_list.txt', accept_list, fmt='%1.2f')
        capi.savetxt(self.path + '/acceptpercent.txt', [accept], fmt='%1.2f')
        return (posterior, fx_train_all, fx_test_all, rmse_train, rmse_test, acc_train, acc_test, likelihood_vec.T,
                accept_vec, accept)

    def make_directory(self, directory):
        if not os.path.exists(directory):
            os.makedirs(directory)


RESULT_DIRS = ['/predictions/', '/posterior', '/results', '/surrogate', '/surrogate/learnsurrogate_data',
               '/posterior/pos_w', '/posterior/pos_likelihood', '/posterior/surg_likelihood',
               '/posterior/accept_list']                                          # R:997
