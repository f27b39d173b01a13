"""Per-launch overhead of the chain kernel on the bench workload: the same 100 MCMC steps (10 swap rounds) as 10 launches
of 10 steps and as 1 launch of 100 steps (CUDA events on the handle's stream)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import load_workload
from ptnn_b200.sampler import Sampler

w = load_workload("synth_ts", 1)
stream = torch.cuda.current_stream()
def make():
    s = Sampler(w["task"], w["topology"], w["temperatures"], 232, 10, use_langevin_gradients=True, l_prob=0.5, learn_rate=0.01,
                seed=2026, common_random_numbers=True, memoize_gradient=0, stream=stream)
    s.set_data(w["train"], w["test"]); s.init_chains(np.random.RandomState(1000).randn(s.R, s.P))
    s.run(30); torch.cuda.synchronize()
    return s
for chunks in (10, 1, 10, 1):
    s = make()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(chunks):
        s.run(100 // chunks)
    b.record(stream); torch.cuda.synchronize()
    print("%2d launch(es) x %3d steps: %.2f ms" % (chunks, 100 // chunks, a.elapsed_time(b)))
    s.close()
