"""Small target for compute-sanitizer (memcheck): a Sunspot ladder through Langevin steps, swap rounds, the left-over
round and speculative windows, then the read-back paths.  usage: sanitizer_target.py [speculation]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptnn_b200.sampler import Sampler, geometric_ladder

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "datasets.npz"))
tr, te = d["reg_Sunspot_train"], d["reg_Sunspot_test"]
spec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
R, S, si = 4, 61, 10
with Sampler(0, (4, 5, 1), geometric_ladder(R, 2), S, si, learn_rate=0.1, l_prob=0.5, seed=5, speculation=spec, debug_traces=True) as s:
    s.set_data(tr, te)
    s.init_chains(np.random.RandomState(1).randn(R, s.P))
    s.run(25)
    s.run()
    t = s.traces()
    sm = s.trace_summary(1, S - 1)
    print("ok", s.step, int(t["accepted"].sum()), s.swap_stats()[:2], float(sm["rmse_train"]["mean"]))
