"""Drop-in for multicore-pt-regression/pt_timeseries_regression.py (R:): same classes, same
constructor / method signatures, same return tuples and output files -- computed on the B200.

    from ptnn_b200.regression import Network, ptReplica, ParallelTempering
    pt = ParallelTempering(use_langevin_gradients, learn_rate, traindata, testdata, topology,
                           num_chains, maxtemp, NumSample, swap_interval, langevin_prob, path)
    pt.initialize_chains(burn_in)
    pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, likelihood_rep, \
        swap_perc, accept_vec, accept = pt.run_chains()
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import _surface as _s
from ._surface import RESULT_DIRS


class Network(_s.NetworkBase):
    """R:27-134."""
    TASK = _s.REGRESSION


class ptReplica(_s.ReplicaBase):
    """R:138-485."""
    TASK = _s.REGRESSION
    NETWORK = Network

    def __init__(self, use_langevin_gradients, learn_rate, w, minlim_param, maxlim_param, samples, traindata,
                 testdata, topology, burn_in, temperature, swap_interval, langevin_prob, path, parameter_queue,
                 main_process, event):
        self._init_common(use_langevin_gradients, learn_rate, w, minlim_param, maxlim_param, samples, traindata,
                          testdata, topology, burn_in, temperature, swap_interval, langevin_prob, path,
                          parameter_queue, main_process, event)

    def likelihood_func(self, fnn, data, w, tau_sq):
        """R:200-205 -> [loglik / adapttemp, fx, rmse]."""
        return self._likelihood(fnn, data, w, tau_sq)

    def prior_likelihood(self, sigma_squared, nu_1, nu_2, w, tausq):
        """R:215-221."""
        return self._prior(sigma_squared, nu_1, nu_2, w, tausq)


class ParallelTempering(_s.ParallelTemperingBase):
    """R:487-875."""
    TASK = _s.REGRESSION
    REPLICA = ptReplica

    def __init__(self, use_langevin_gradients, learn_rate, traindata, testdata, topology, num_chains, maxtemp,
                 NumSample, swap_interval, langevin_prob, path):
        self._init_common(use_langevin_gradients, learn_rate, traindata, testdata, topology, num_chains, maxtemp,
                          NumSample, swap_interval, langevin_prob, path)

    def _make_replica(self, w, i):                                               # R:650
        return ptReplica(self.use_langevin_gradients, self.learn_rate, w, self.minlim_param, self.maxlim_param,
                         self.NumSamples, self.traindata, self.testdata, self.topology, self.burn_in,
                         self.temperatures[i], self.swap_interval, self.langevin_prob, self.path,
                         self.parameter_queue[i], self.wait_chain[i], self.event[i])


PROBLEMS = {1: "Lazer", 2: "Sunspot", 3: "Mackey", 4: "Lorenz", 5: "Rossler", 6: "Henon", 7: "ACFinance"}   # R:882-909


def run_problem(problem, data_root, out_root, *, hidden=5, NumSample=100000, maxtemp=2, swap_ratio=0.01,
                num_chains=10, burn_in=0.5, learn_rate=0.1, use_langevin_gradients=True, langevin_prob=0.5,
                seed=None, results="host"):
    """One iteration of the reference's main() loop (R:879-1061) without the plots: runs the
    sampler and appends the 15-number row to master_result_file.txt / result.txt.
    results="host": run_chains() as the reference does (per-chain txt files, the 11-tuple);
    results="device": the statistics of the row are reduced on the device traces (run_summary)."""
    name = PROBLEMS[problem]
    traindata = np.loadtxt(os.path.join(data_root, "Data_OneStepAhead", name, "train.txt"))
    testdata = np.loadtxt(os.path.join(data_root, "Data_OneStepAhead", name, "test.txt"))
    topology = [4, hidden, 1]
    swap_interval = int(swap_ratio * NumSample / num_chains)                     # R:949
    run_nb = 0
    while os.path.exists(os.path.join(out_root, name + '_%s' % run_nb)):
        run_nb += 1
    path = os.path.join(out_root, name + '_%s' % run_nb)
    os.makedirs(path)
    timer = time.time()
    pt = ParallelTempering(use_langevin_gradients, learn_rate, traindata, testdata, topology, num_chains, maxtemp,
                           NumSample, swap_interval, langevin_prob, path)
    pt.seed = seed
    for d in RESULT_DIRS:
        pt.make_directory(path + d)
    pt.initialize_chains(burn_in)
    if results == "device":
        sm = pt.run_summary()
        tr, te, swap_perc, accept_per = sm["rmse_train"], sm["rmse_test"], sm["swap_perc"], sm["accept_per"]
        stats = [tr["mean"], tr["std"], tr["min"], te["mean"], te["std"], te["min"]]
    else:
        (pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, likelihood_rep, swap_perc, accept_vec,
         accept) = pt.run_chains()
        list_end = accept_vec.shape[1]
        accept_ratio = accept_vec[:, list_end - 1:list_end] / list_end           # R:1009-1011 (Q16)
        accept_per = np.mean(accept_ratio) * 100
        stats = [np.mean(rmse_train), np.std(rmse_train), np.amin(rmse_train),
                 np.mean(rmse_test), np.std(rmse_test), np.amin(rmse_test)]
    timetotal = (time.time() - timer) / 60
    allres = np.asarray([problem, NumSample, maxtemp, swap_interval, langevin_prob, learn_rate] + stats +
                        [swap_perc, accept_per, timetotal])                     # R:1052
    xv = name + '_' + str(run_nb)
    for fn in (os.path.join(path, 'result.txt'), os.path.join(out_root, 'master_result_file.txt')):
        with open(fn, "a+") as f:
            np.savetxt(f, allres, fmt='%1.4f', newline=' ')
            np.savetxt(f, [xv], fmt="%s", newline=' \n')
    return allres, pt


def main(argv=None):
    """python -m ptnn_b200.regression [problem ...]  (env PT_DATA_ROOT = directory holding
    Data_OneStepAhead/, PT_OUT_ROOT = results directory; the reference hard-codes both, R:883, R:960)."""
    argv = sys.argv[1:] if argv is None else argv
    data_root = os.environ.get("PT_DATA_ROOT", ".")
    out_root = os.environ.get("PT_OUT_ROOT", "Res_LG-Lprob")
    os.makedirs(out_root, exist_ok=True)
    for p in ([int(a) for a in argv] or [1]):
        allres, _ = run_problem(p, data_root, out_root)
        print(PROBLEMS[p], allres)


if __name__ == "__main__":
    main()
