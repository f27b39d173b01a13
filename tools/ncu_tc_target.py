"""Short target for ncu: the tcgen05 likelihood kernel (op_forward_tc_kernel) on the PenDigit-shaped set."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptnn_b200 import capi, datasets
from oracle import ptfnn_numpy as on
tr, te = datasets.synthetic_pendigit()
topo = (16, 256, 10)
w = np.random.RandomState(0).randn(on.num_params(topo)) * 0.3
capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, tr, w, 1.0, 1.0)
t0 = time.perf_counter()
out = capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, tr, w, 1.0, 1.0)
print("ok", out[:3], "wall %.2f ms (includes packing and copies)" % ((time.perf_counter() - t0) * 1e3))
