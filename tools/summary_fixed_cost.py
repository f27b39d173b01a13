"""Fixed cost vs streaming rate of trace_summary_kernel: time it over traces of growing length with the L2 evicted by
a write (dirty lines), by a read (clean lines) and not at all (tools; timing only, values are whatever the buffers hold)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptnn_b200.sampler import Sampler, geometric_ladder
from oracle import ptfnn_numpy as on

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
topo, R = (4, 64, 1), 1024
P = on.num_params(topo)
rs = np.random.RandomState(0)
tr = rs.rand(64, 5)
for S in (66, 131, 261, 521, 1041):
    with Sampler(on.REGRESSION, topo, geometric_ladder(R, 2), S, 10) as s:
        s.set_data(tr, tr)
        s.init_chains(rs.randn(R, P))
        s.run(10)
        out = []
        for mode in ("write", "read", "none"):
            ms = []
            for _ in range(5):
                if mode == "write":
                    flush.fill_(1)
                elif mode == "read":
                    flush.sum()
                torch.cuda.synchronize()
                ms.append(s.trace_summary(1, S - 1)["kernel_ms"])
            out.append(min(ms[1:]))
        b = R * (S - 1) * (32 + 4 * P)
        print("rows %4d  %7.1f MB   evict-by-write %.1f us   evict-by-read %.1f us   no eviction %.1f us   (ideal at 6545 GB/s: %.1f us)" %
              (S - 1, b / 1e6, out[0] * 1e3, out[1] * 1e3, out[2] * 1e3, b / 6545e3))
