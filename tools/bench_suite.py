"""Throughput on the reference's own data sets (BASELINE.json configs[1] and configs[2]): the seven
Data_OneStepAhead series with FNN 4-5-1 (random-walk and Langevin proposals, swap interval 100) and the
three classification sets (Iris 4-12-3, Ionosphere 34-50-2, Cancer 9-12-2; C:1036-1045), 10 temperatures each.
Device-resident replica-steps/s (CUDA events, free-running Philox draws), one JSON line per case."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptnn_b200.sampler import Sampler, geometric_ladder

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "datasets.npz"))
cases = []
for name in ("Lazer", "Sunspot", "Mackey", "Lorenz", "Rossler", "Henon", "ACFinance"):
    for lg in (False, True):
        cases.append(dict(name="reg_" + name, task=0, topo=(4, 5, 1), maxtemp=2, si=100, lr=0.1, lg=lg))
for name in ("Iris", "Ionosphere", "Cancer"):
    topo = tuple(int(x) for x in d["cls_%s_topology" % name])
    for lg in (False, True):
        cases.append(dict(name="cls_" + name, task=1, topo=topo, maxtemp=10, si=100, lr=0.01, lg=lg))
R, n_launch = 10, 6
for c in cases:
    tr, te = d[c["name"] + "_train"], d[c["name"] + "_test"]
    S = c["si"] * (n_launch + 2) + 3
    for memo in (0, 1):
        s = Sampler(c["task"], c["topo"], geometric_ladder(R, c["maxtemp"]), S, c["si"], use_langevin_gradients=c["lg"],
                    l_prob=0.5, learn_rate=c["lr"], memoize_gradient=memo, seed=1, stream=torch.cuda.current_stream())
        s.set_data(tr, te)
        s.init_chains(np.random.RandomState(0).randn(R, s.P))
        s.run(2 * c["si"] + (1 if c["task"] == 0 else 0))      # launches end on swap rounds (R:427 | C:438)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(n_launch):
            s.run(c["si"])
        b.record(); torch.cuda.synchronize()
        c["steps_per_s_memo%d" % memo] = R * c["si"] * n_launch / (a.elapsed_time(b) * 1e-3)
        s.close()
    print(json.dumps(dict(dataset=c["name"], topology=c["topo"], rows=[int(tr.shape[0]), int(te.shape[0])], proposals="langevin" if c["lg"] else "random-walk",
                          replica_steps_per_s=round(c["steps_per_s_memo0"]), replica_steps_per_s_memoized=round(c["steps_per_s_memo1"]))))
