"""Measurement helper: where the time of a small-ladder run with speculative windows goes.  Runs the Sunspot
configuration of the bench, then replays the window plan of chain_body (32 steps at most, one Langevin step per CTA,
"apart" or "riding") on the recorded acceptances and lx draws, and compares the
number of windows of the slowest temperature per swap segment with the measured time."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptnn_b200.sampler import Sampler, geometric_ladder

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "datasets.npz"))
tr, te = d["reg_Sunspot_train"], d["reg_Sunspot_test"]
R, S, si, K = 10, 4002, 50, 14
s = Sampler(0, (4, 5, 1), geometric_ladder(R, 2), S, si, learn_rate=0.1, l_prob=0.5, seed=2026, memoize_gradient=0,
            debug_traces=True, stream=torch.cuda.current_stream())
s.set_data(tr, te)
s.init_chains(np.random.RandomState(1000).randn(R, s.P))
s.run(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); s.run(4000); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
acc = s.traces(pos_w=False)["accepted"][:, 2:4002].astype(bool)        # row i+1 = step i; steps 1..4000
lx = s.generate_draws(1, 4000)[0]
lg = lx < 0.5
nseg = 4000 // si
wmax, wsum, lgmax = [], [], []
for g in range(nseg):
    per = []
    for r in range(R):
        L, A = lg[r, g * si:(g + 1) * si], acc[r, g * si:(g + 1) * si]
        i, nw = 0, 0
        while i < si:
            # chain_body's plan: "apart" (Langevin step j on CTA j, random-walk steps eight to a CTA from the top) unless
            # "riding" (owner = Langevin steps before the step) covers more than two steps more
            wcap = min(32, si - i)
            n_lg = m = w_apart = w_riding = 0
            for t in range(wcap):
                if max(n_lg, t // 9) >= K: break
                n_lg += int(L[i + t]); w_riding += 1
            n_lg = 0
            for t in range(wcap):
                n2, m2 = n_lg + int(L[i + t]), m + int(not L[i + t])
                if n2 + (m2 + 7) // 8 > K: break
                n_lg, m, w_apart = n2, m2, w_apart + 1
            w = w_apart if w_apart + (2 if K >= 8 else 0) >= w_riding else w_riding
            hit = np.flatnonzero(A[i:i + w])
            i += (hit[0] + 1) if hit.size else w
            nw += 1
        per.append(nw)
    wmax.append(max(per)); wsum.append(np.mean(per))
print("Sunspot R=10 K=14: %.1f us per 50-step segment; windows per segment: slowest temperature %.2f, mean temperature %.2f; acceptance %.3f (hottest %.3f)" % (
    1e3 * ms / nseg, np.mean(wmax), np.mean(wsum), acc.mean(), acc.mean(axis=1).max()))
print("=> %.1f us per window of the slowest temperature (a Langevin step is ~54 us, a random-walk step ~4 us)" % (1e3 * ms / nseg / np.mean(wmax)))
s.close()
