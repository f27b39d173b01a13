"""Drop-in for multicore-pt-classification/pt_classification.py (C:): same classes, signatures,
return tuples and output files -- computed on the B200.  ``python -m ptnn_b200.classification
<swap_ratio>`` honours the run.sh argument (run.sh:8-11) that the reference parses at C:1039.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import _surface as _s
from ._surface import RESULT_DIRS


class Network(_s.NetworkBase):
    """C:26-153."""
    TASK = _s.CLASSIFICATION

    def softmax(self):
        prob = np.exp(self.out) / np.sum(np.exp(self.out))                      # C:108-110
        return prob


class ptReplica(_s.ReplicaBase):
    """C:157-494 (no langevin_prob argument: l_prob is fixed at 0.5, C:192)."""
    TASK = _s.CLASSIFICATION
    NETWORK = Network

    def __init__(self, use_langevin_gradients, learn_rate, w, minlim_param, maxlim_param, samples, traindata,
                 testdata, topology, burn_in, temperature, swap_interval, path, parameter_queue, main_process, event):
        self._init_common(use_langevin_gradients, learn_rate, w, minlim_param, maxlim_param, samples, traindata,
                          testdata, topology, burn_in, temperature, swap_interval, 0.5, path, parameter_queue,
                          main_process, event)

    def likelihood_func(self, fnn, data, w):
        """C:209-222 -> [loglik / adapttemp, fx (argmax), rmse]."""
        return self._likelihood(fnn, data, w, 1.0)

    def prior_likelihood(self, sigma_squared, nu_1, nu_2, w):
        """C:224-230."""
        return self._prior(sigma_squared, nu_1, nu_2, w, 1.0)


class ParallelTempering(_s.ParallelTemperingBase):
    """C:497-897."""
    TASK = _s.CLASSIFICATION
    REPLICA = ptReplica

    def __init__(self, use_langevin_gradients, learn_rate, traindata, testdata, topology, num_chains, maxtemp,
                 NumSample, swap_interval, path):
        self._init_common(use_langevin_gradients, learn_rate, traindata, testdata, topology, num_chains, maxtemp,
                          NumSample, swap_interval, 0.5, path)

    def _make_replica(self, w, i):                                               # C:659
        return ptReplica(self.use_langevin_gradients, self.learn_rate, w, self.minlim_param, self.maxlim_param,
                         self.NumSamples, self.traindata, self.testdata, self.topology, self.burn_in,
                         self.temperatures[i], self.swap_interval, self.path, self.parameter_queue[i],
                         self.wait_chain[i], self.event[i])


def _zscore_and_split(features, classes, train_ratio=0.7, complementary_test_split=False):
    """C:1001-1012 (``separate_flag``): z-score every feature column, then a random split drawn from the
    global NumPy stream (np.random.seed() before the call makes it repeatable, as in the reference).
    The training set is the first 70 % of a random permutation.  The reference's TEST set is
    ``features[indices[k]:, :]`` with k = int(0.7 n) (C:1011-1012: the bracket closes after the index, so the
    permutation's k-th ENTRY is used as the start of a plain slice): every row from that row number to the end
    of the file, overlapping the training rows -- reproduced here (DESIGN quirk Q17).
    ``complementary_test_split=True`` gives the split the line was presumably meant to be (the other 30 %)."""
    features = np.array(features, dtype=np.float64)
    for k in range(features.shape[1]):
        features[:, k] = (features[:, k] - np.mean(features[:, k])) / np.std(features[:, k])
    indices = np.random.permutation(features.shape[0])                            # C:1010
    ntr = int(train_ratio * features.shape[0])
    traindata = np.hstack([features[indices[:ntr], :], classes[indices[:ntr], :]])
    if complementary_test_split:
        testdata = np.hstack([features[indices[ntr:], :], classes[indices[ntr:], :]])
    else:
        testdata = np.hstack([features[indices[ntr]:, :], classes[indices[ntr]:, :]])
    return traindata, testdata


def load_problem(problem, data_root, complementary_test_split=False):
    """The dataset switch of C:909-1012 for the problems whose files ship with the reference
    (6 Bank needs DATA/Bank/bank-processed.csv and 8 Chess DATA/chess.csv, which the repository does not hold)."""
    base = os.path.join(data_root, "DATA")
    if problem in (1, 2):                                                         # wine quality red / white, C:909-919, C:931-941
        name = "winequality-red" if problem == 1 else "winequality-white"
        data = np.genfromtxt(os.path.join(base, name + '.csv'), delimiter=';')[1:, :]     # first line: column labels
        tr, te = _zscore_and_split(data[:, 0:11], data[:, 11].reshape(data.shape[0], 1),  # quality scores 3..9 index 10 outputs as they are
                                   complementary_test_split=complementary_test_split)
        return name, tr, te, [11, 50, 10]
    if problem == 3:                                                              # Iris, C:920-930
        data = np.genfromtxt(os.path.join(base, 'iris.csv'), delimiter=';')
        tr, te = _zscore_and_split(data[:, 0:4], data[:, 4].reshape(data.shape[0], 1) - 1,
                                   complementary_test_split=complementary_test_split)
        return "iris", tr, te, [4, 12, 3]
    if problem == 4:                                                              # Ionosphere, C:942-949
        tr = np.genfromtxt(os.path.join(base, 'Ions/Ions/ftrain.csv'), delimiter=',')[:, :-1]
        te = np.genfromtxt(os.path.join(base, 'Ions/Ions/ftest.csv'), delimiter=',')[:, :-1]
        return "Ionosphere", tr, te, [34, 50, 2]
    if problem == 5:                                                              # Cancer, C:950-957
        tr = np.genfromtxt(os.path.join(base, 'Cancer/ftrain.txt'), delimiter=' ')[:, :-1]
        te = np.genfromtxt(os.path.join(base, 'Cancer/ftest.txt'), delimiter=' ')[:, :-1]
        return "Cancer", tr, te, [9, 12, 2]
    if problem == 7:                                                              # PenDigit, C:972-986
        tr = np.genfromtxt(os.path.join(base, 'PenDigit/train.csv'), delimiter=',')
        te = np.genfromtxt(os.path.join(base, 'PenDigit/test.csv'), delimiter=',')
        for k in range(16):
            tr[:, k] = (tr[:, k] - np.mean(tr[:, k])) / np.std(tr[:, k])
            te[:, k] = (te[:, k] - np.mean(te[:, k])) / np.std(te[:, k])
        return "PenDigit", tr, te, [16, 30, 10]
    raise ValueError("problem %r: dataset not shipped with the reference (.MISSING_LARGE_BLOBS) or unknown" % problem)


def run_problem(problem, data_root, out_root, *, NumSample=50000, maxtemp=10, swap_ratio=0.02, num_chains=10,
                burn_in=0.5, learn_rate=0.01, use_langevin_gradients=False, seed=None, results="host"):
    """One iteration of the reference's main() loop (C:901-1147) without the plots.
    results="device": the statistics of the row are reduced on the device traces (run_summary)."""
    name, traindata, testdata, topology = load_problem(problem, data_root)
    swap_interval = int(swap_ratio * (NumSample / num_chains))                   # C:1045
    run_nb = 0
    while os.path.exists(os.path.join(out_root, name + '_%s' % run_nb)):
        run_nb += 1
    path = os.path.join(out_root, name + '_%s' % run_nb)
    os.makedirs(path)
    timer = time.time()
    pt = ParallelTempering(use_langevin_gradients, learn_rate, traindata, testdata, topology, num_chains, maxtemp,
                           NumSample, swap_interval, path)
    pt.seed = seed
    for d in RESULT_DIRS:
        pt.make_directory(path + d)
    pt.initialize_chains(burn_in)
    if results == "device":
        sm = pt.run_summary()
        tr, te, swap_perc, accept_per = sm["acc_train"], sm["acc_test"], sm["swap_perc"], sm["accept_per"]
        stats = [tr["mean"], tr["std"], tr["max"], te["mean"], te["std"], te["max"]]
    else:
        (pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, likelihood_rep, swap_perc, accept_vec,
         accept) = pt.run_chains()
        list_end = accept_vec.shape[1]
        accept_ratio = accept_vec[:, list_end - 1:list_end] / list_end
        accept_per = np.mean(accept_ratio) * 100
        stats = [np.mean(acc_train), np.std(acc_train), np.amax(acc_train),
                 np.mean(acc_test), np.std(acc_test), np.amax(acc_test)]
    timetotal = (time.time() - timer) / 60
    allres = np.asarray([problem, NumSample, maxtemp, swap_interval, use_langevin_gradients, learn_rate] + stats +
                        [swap_perc, accept_per, timetotal])                     # C:1138
    xv = name + '_' + str(run_nb)
    for fn in (os.path.join(path, 'result.txt'), os.path.join(out_root, 'master_result_file.txt')):
        with open(fn, "a+") as f:
            np.savetxt(f, allres, fmt='%1.2f', newline=' ')                    # C:1140-1146
            np.savetxt(f, [xv], fmt="%s", newline=' \n')
    return allres, pt


def main(argv=None):
    """python -m ptnn_b200.classification <swap_ratio> [problem ...]   (run.sh:8-11 passes swap_ratio)."""
    argv = sys.argv[1:] if argv is None else argv
    swap_ratio = float(argv[0]) if argv else 0.02                                # C:1039
    data_root = os.environ.get("PT_DATA_ROOT", ".")
    out_root = os.environ.get("PT_OUT_ROOT", "PT_EvalSwap")
    os.makedirs(out_root, exist_ok=True)
    for p in ([int(a) for a in argv[1:]] or [3, 4, 5]):
        allres, _ = run_problem(p, data_root, out_root, swap_ratio=swap_ratio)
        print(p, allres)


if __name__ == "__main__":
    main()
