"""Measurement helper (GPU box): Langevin / random-walk step cost of the synth_ts chain kernel as a
function of the number of co-resident temperatures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_perf import chain_step_cost
from ptnn_b200 import datasets
tr, te = datasets.synthetic_timeseries()
for R in [int(a) for a in sys.argv[1:]] or [148, 296, 592, 888, 1024]:
    c0 = chain_step_cost(0, (4, 64, 1), tr, te, R, 10, 0.01, 0, n=4)
    c1 = chain_step_cost(0, (4, 64, 1), tr, te, R, 10, 0.01, 1, n=4)
    p1 = c0["LG"] - c1["LG"]
    print("R=%4d  LG memo0 %.3f ms  memo1 %.3f ms  (pass-1 %.3f ms = %.0f cycles/row)  RW %.3f ms" % (
        R, c0["LG"], c1["LG"], p1, p1 * 1e-3 * 1.965e9 / tr.shape[0], c0["RW"]))
