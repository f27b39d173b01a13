"""Short target for ncu: 128 of the 4-64-1 temperatures (the strong-scaling split of configs[3] at 8 GPUs) with eight
CTAs per temperature, burned in, then three launches of one swap segment each.  Profile the last launches:
ncu -k regex:chain_kernel -s <launches of the burn-in> -c 2 ...   usage: ncu_spec_target.py [burn steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptnn_b200 import datasets
from ptnn_b200.sampler import Sampler, geometric_ladder

burn = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
tr, te = datasets.synthetic_timeseries()
R = 128
s = Sampler(0, (4, 64, 1), geometric_ladder(R, 2), burn + 40, 10, learn_rate=0.01, l_prob=0.5, seed=2026, memoize_gradient=0, speculation=8)
s.set_data(tr, te)
s.init_chains(np.random.RandomState(1000).randn(R, s.P))
s.run(burn + 1)            # one launch (a fixed depth is not launched in pieces)
for _ in range(3):
    s.run(10)
s.sync()
print("ok", s.step, float(s.get_state()["num_accepted"].mean()))
