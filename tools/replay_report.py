import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from oracle import ptfnn_c as oc
from tests import common as cm
from tests.test_gpu_replay import _sampler, _first_divergence
for name in cm.CASES:
    fx, cfg, tr, te, draws = cm.case(name)
    ref = oc.run_pt(cfg, tr, te, fx["temperatures"], fx["w0"], draws)
    with _sampler(cfg, fx["temperatures"]) as s:
        s.set_data(tr, te); s.init_chains(fx["w0"]); s.replay(draws)
        t = s.traces(); ns, tot, sw = s.swap_stats()
    i_star, i_acc, i_sw = _first_divergence(t, ref, sw, cfg)
    rows = slice(0, i_star + 1)
    print(name, "S-1 =", cfg.samples - 1, "i_star =", i_star, "lik relerr %.2e pos_w %.2e" % (cm.relerr(t["lik_prop"][:, rows][:, 1:], ref.lik_prop[:, rows][:, 1:]), cm.relerr(t["pos_w"][:, rows], fx["ref_pos_w"][:, rows])))
