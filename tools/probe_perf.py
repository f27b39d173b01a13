"""Measurement helper (GPU box): per-row latency of the serial SGD recurrence and per-step cost of
Langevin vs random-walk steps of the chain kernel.  Not part of the product or the tests."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptnn_b200 import capi, datasets
from ptnn_b200.sampler import Sampler, geometric_ladder
from oracle import ptfnn_numpy as on


def sgd_row_latency(task, topo, data, lr=0.01, d0=1, d1=9):
    w = np.random.RandomState(0).randn(on.num_params(topo)) * 0.3
    ms = capi.time_langevin_gradient(task, topo, data, w, lr, depth=1, repeats=3)
    return ms * 1e6 / data.shape[0]


def chain_step_cost(task, topo, train, test, R, si, lr, memo, n=8):
    """ms per PT step of the whole ladder, all-Langevin (l_prob = 1) and all-random-walk (l_prob = 0),
    free-running Philox draws (nothing is uploaded inside the timed region), CUDA events."""
    temps = geometric_ladder(R, 2) if R > 1 else np.ones(1)
    out = {}
    for label, lp in (("LG", 1.0), ("RW", 0.0)):
        S = n * 3 + 2
        s = Sampler(task, topo, temps, S, 10 ** 6, learn_rate=lr, l_prob=lp, memoize_gradient=memo,
                    stream=torch.cuda.current_stream())
        s.set_data(train, test)
        s.init_chains(np.random.RandomState(1).randn(R, s.P) * 0.5)
        s.run(n)                                    # warm-up
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); s.run(n); b.record(); torch.cuda.synchronize()
        out[label] = a.elapsed_time(b) / n
        s.close()
    return out


if __name__ == "__main__":
    print(capi.build_info())
    tr, te = datasets.synthetic_timeseries()
    d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "datasets.npz"))
    sun_tr, sun_te = d["reg_Sunspot_train"], d["reg_Sunspot_test"]
    clk = 1.965
    for name, task, topo, data in (("4-5-1 sunspot-rows x100", 0, (4, 5, 1), np.tile(sun_tr, (100, 1))),
                                   ("4-10-1", 0, (4, 10, 1), np.tile(sun_tr, (100, 1))),
                                   ("4-64-1 synth", 0, (4, 64, 1), tr)):
        ns = sgd_row_latency(task, topo, data)
        print("SGD row latency %-28s %7.1f ns/row = %6.0f cycles @%.3f GHz" % (name, ns, ns * clk, clk))
    ir = np.hstack([np.random.RandomState(2).randn(20000, 16), np.random.RandomState(3).randint(0, 10, (20000, 1)).astype(float)])
    for name, topo in (("16-30-10", (16, 30, 10)), ("16-256-10", (16, 256, 10))):
        ns = sgd_row_latency(1, topo, ir, d1=3)
        print("SGD row latency %-28s %7.1f ns/row = %6.0f cycles" % (name, ns, ns * clk))
    for memo in (0, 1):
        c = chain_step_cost(0, (4, 64, 1), tr, te, 1024, 10, 0.01, memo, n=6)
        print("synth_ts R=1024 memo=%d: LG step %.3f ms, RW step %.3f ms" % (memo, c["LG"], c["RW"]))
        c = chain_step_cost(0, (4, 5, 1), sun_tr, sun_te, 10, 50, 0.1, memo, n=50)
        print("sunspot  R=10   memo=%d: LG step %.4f ms, RW step %.4f ms" % (memo, c["LG"], c["RW"]))
    c = chain_step_cost(0, (4, 64, 1), tr, te, 128, 10, 0.01, 0, n=6)
    print("synth_ts R=128 memo=0: LG step %.3f ms, RW step %.3f ms" % (c["LG"], c["RW"]))
