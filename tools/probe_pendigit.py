"""Measurement helper (GPU box): Langevin / random-walk step cost of the 16-256-10 chain kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from probe_perf import chain_step_cost, sgd_row_latency
from ptnn_b200 import datasets
tr, te = datasets.synthetic_pendigit()
print("rows", tr.shape, te.shape)
ns = sgd_row_latency(1, (16, 256, 10), tr)
print("SGD team row latency 16-256-10: %.1f ns/row = %.0f cycles" % (ns, ns * 1.965))
for R in [int(a) for a in sys.argv[1:]] or [148, 256]:
    c0 = chain_step_cost(1, (16, 256, 10), tr, te, R, 10, 0.01, 0, n=2)
    c1 = chain_step_cost(1, (16, 256, 10), tr, te, R, 10, 0.01, 1, n=2)
    print("R=%4d  LG memo0 %.3f ms  memo1 %.3f ms  RW %.3f ms" % (R, c0["LG"], c1["LG"], c0["RW"]))
