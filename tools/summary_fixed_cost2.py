import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from ptnn_b200.sampler import Sampler, geometric_ladder
from oracle import ptfnn_numpy as on
topo, R, S = (4, 64, 1), 1024, 1041
P = on.num_params(topo)
rs = np.random.RandomState(0)
tr = rs.rand(64, 5)
with Sampler(on.REGRESSION, topo, geometric_ladder(R, 2), S, 10) as s:
    s.set_data(tr, tr); s.init_chains(rs.randn(R, P)); s.run(10)
    def t(first, count, post):
        ms = []
        for _ in range(6):
            torch.cuda.synchronize()
            ms.append(s.trace_summary(first, count, posterior=post)["kernel_ms"])
        return min(ms[1:]) * 1e3
    for count in (1, 2, 8, 21, 42, 65, 1040):
        print("count %4d: moments+series %.1f us, series only %.1f us" % (count, t(1, count, True), t(1, count, False)))
