"""Short target for ncu: the single-warp SGD recurrence kernel (op_sgd_kernel) on the synthetic series.
usage: ncu_sgd_target.py [H]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptnn_b200 import capi, datasets
from oracle import ptfnn_numpy as on

tr, te = datasets.synthetic_timeseries()
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
topo = (4, H, 1)
w = np.random.RandomState(0).randn(on.num_params(topo)) * 0.3
ms = capi.time_langevin_gradient(0, topo, tr, w, 0.01, depth=1, repeats=2)
print("ok %.3f ms, %.1f ns/row" % (ms, ms * 1e6 / tr.shape[0]))
