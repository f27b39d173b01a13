"""Short target for ncu / timing: trace_summary_kernel (result pipeline on the device traces)
on traces of the synthetic-series shape (R temperatures x S rows x P = 385) and of a small net (P = 31).
The chains are only initialised (the values reduced are whatever the trace buffers hold): timing only."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptnn_b200.sampler import Sampler, geometric_ladder
from oracle import ptfnn_numpy as on

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for topo, R, S in (((4, 64, 1), 1024, int(os.environ.get("PT_S", 1001))), ((4, 5, 1), 1024, 4001), ((16, 256, 10), 64, 501)):
    task = on.REGRESSION if topo[2] == 1 else on.CLASSIFICATION
    P = on.num_params(topo)
    rs = np.random.RandomState(0)
    tr = rs.rand(64, topo[0] + 1)
    if task == on.CLASSIFICATION:
        tr[:, -1] = rs.randint(0, topo[2], 64)
    with Sampler(task, topo, geometric_ladder(R, 2), S, 10) as s:
        s.set_data(tr, tr)
        s.init_chains(rs.randn(R, P))
        s.run(10)
        ms = []
        for _ in range(4):
            flush.fill_(1); torch.cuda.synchronize()
            ms.append(s.trace_summary(1, S - 1)["kernel_ms"])
        b = R * (S - 1) * (32 + 4 * P)
        print("topology %s R %d rows %d: %.1f MB, kernel %.3f ms (min %.3f) -> %.0f GB/s" %
              (topo, R, S - 1, b / 1e6, np.mean(ms[1:]), min(ms), b / min(ms) / 1e6))
