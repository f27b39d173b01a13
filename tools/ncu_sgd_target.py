"""Short target for ncu: the SGD recurrence kernel (op_sgd_kernel) alone.
usage: ncu_sgd_target.py [synth H | pendigit]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptnn_b200 import capi, datasets
from oracle import ptfnn_numpy as on

if len(sys.argv) > 1 and sys.argv[1] == "pendigit":
    tr, te = datasets.synthetic_pendigit()
    task, topo = 1, (16, 256, 10)
else:
    tr, te = datasets.synthetic_timeseries()
    task, topo = 0, (4, int(sys.argv[2]) if len(sys.argv) > 2 else 64, 1)
w = np.random.RandomState(0).randn(on.num_params(topo)) * 0.3
ms = capi.time_langevin_gradient(task, topo, tr, w, 0.01, depth=1, repeats=2)
print("ok %.3f ms, %.1f ns/row" % (ms, ms * 1e6 / tr.shape[0]))
