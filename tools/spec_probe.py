"""Measurement helper: ms per 10-step launch of the 4-64-1 ladder at R temperatures for fixed and automatic window depths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ptnn_b200 import datasets
from ptnn_b200.sampler import Sampler, geometric_ladder
tr, te = datasets.synthetic_timeseries()
BURN = int(os.environ.get("PROBE_BURN", "0"))          # steps run (untimed) before the measurement
for R in [int(a) for a in sys.argv[1:]] or [128]:
    for spec in (1, 0, 4, 8):
        n = 30
        s = Sampler(0, (4, 64, 1), geometric_ladder(R, 2), 10 * n + 3 + BURN, 10, learn_rate=0.01, l_prob=0.5, seed=2026, memoize_gradient=0,
                    stream=torch.cuda.current_stream(), speculation=spec)
        s.set_data(tr, te)
        s.init_chains(np.random.RandomState(1000).randn(R, s.P))
        s.run(BURN + 1)            # + 1: launches of 10 steps end on the swap rounds (R:427)
        acc0 = s.get_state()["num_accepted"].sum()
        ev = []
        for k in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.run(10); b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        acc = (s.get_state()["num_accepted"].sum() - acc0) / (R * 10.0 * n)
        print("R %4d speculation %d: ms per launch first 5 %s ... last 5 %s  mean(last 20) %.2f  acceptance %.3f" % (
            R, spec, np.round(ms[:5], 1), np.round(ms[-5:], 1), np.mean(ms[-20:]), acc))
        s.close()
