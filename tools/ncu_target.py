"""Short target for ncu: a few launches of the chain kernel on the synth_ts / sunspot workloads with
forced all-random-walk and all-Langevin steps.  usage: ncu_target.py [synth_ts|sunspot] [R] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptnn_b200 import datasets
from ptnn_b200.sampler import Sampler, geometric_ladder
from oracle import ptfnn_numpy as on

wl = sys.argv[1] if len(sys.argv) > 1 else "synth_ts"
task = 0
if wl == "pendigit":
    tr, te = datasets.synthetic_pendigit(); topo, R, lr, n, task = (16, 256, 10), 256, 0.01, 1, 1
elif wl == "synth_ts":
    tr, te = datasets.synthetic_timeseries(); topo, R, lr, n = (4, 64, 1), 1024, 0.01, 2
else:
    d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "datasets.npz"))
    tr, te = d["reg_Sunspot_train"], d["reg_Sunspot_test"]; topo, R, lr, n = (4, 5, 1), 10, 0.1, 50
R = int(sys.argv[2]) if len(sys.argv) > 2 else R
n = int(sys.argv[3]) if len(sys.argv) > 3 else n
S = 4 * n + 2
s = Sampler(task, topo, geometric_ladder(R, 2), S, 10 ** 6, learn_rate=lr, memoize_gradient=0)
s.set_data(tr, te)
s.init_chains(np.random.RandomState(1).randn(R, s.P))
lx, z, ze, u = s.generate_draws(0, S - 1)
for k, v in enumerate((0.999, 0.0, 0.999, 0.0)):     # launches: RW, LG, RW, LG  (n steps each)
    lx[:, k * n:(k + 1) * n] = v
d = on.Draws(lx=lx, z=z, z_eta=ze, u=u, u_swap=None)
for k in range(4):
    s.replay(d, n_steps=n)
s.sync()
print("ok", s.step)
