"""Worker for tests/test_gpu_distributed.py (torch.distributed.run, NCCL, one rank per GPU):
the ladder split over ranks must give traces bit-identical to the single-GPU run of the same
replay (tier 4)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ptfnn_numpy as on                      # noqa: E402
from ptnn_b200.distributed import make_gpu_ladder, partition    # noqa: E402
from ptnn_b200.sampler import Sampler, geometric_ladder    # noqa: E402
from tests import common as cm                            # noqa: E402


def main():
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    mode = os.environ.get("PT_TEST_MODE", "replay")
    peer = os.environ.get("PT_TEST_EXCHANGE", "peer") == "peer"
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    Rg, S, si = 8 * world, 62, 5
    cfg = on.PTConfig(task=on.REGRESSION, topology=(4, 5, 1), samples=S, swap_interval=si,
                      use_langevin_gradients=True, l_prob=0.5, learn_rate=0.1)
    temps = geometric_ladder(Rg, 2)
    w0 = np.random.RandomState(11).randn(Rg, cfg.P)
    draws = on.random_draws(cfg, Rg, seed=5, common_random_numbers=False)
    draws.u_swap[:] = draws.u_swap * 0.3                  # frequent swaps: rows really cross rank boundaries
    lo, n = partition(Rg, world, rank)
    kw = dict(use_langevin_gradients=True, l_prob=0.5, learn_rate=0.1, seed=77, common_random_numbers=False)
    spec = int(os.environ.get("PT_TEST_SPEC", "1"))           # speculative windows on every rank (the single-GPU comparison runs without)
    ladder, smp = make_gpu_ladder(on.REGRESSION, (4, 5, 1), temps, S, si, device=local_rank, debug_traces=True, peer=peer,
                                  speculation=spec, **kw)
    smp.set_data(tr, te)

    def one_pass(w0, tag):
        ladder.init_chains(w0[lo:lo + n])      # (re-)initialises the same handles: the peer flag domain moves on
        if mode == "replay":
            local = on.Draws(lx=draws.lx[lo:lo + n], z=draws.z[lo:lo + n], z_eta=draws.z_eta[lo:lo + n],
                             u=draws.u[lo:lo + n], u_swap=None)
            ladder.run(None, local, draws.u_swap)
        else:
            ladder.run(None)
        t = smp.traces()
        ns, tot, sw = smp.swap_stats()
        pooled = ladder.summary(S // 2, S - S // 2)              # collective: every rank
        st = smp.get_state()
        pack = np.concatenate([t["pos_w"].reshape(n, -1), t["lik_prop"], t["accept_list"], st["w"], st["eta"][:, None]], axis=1)
        mine = torch.from_numpy(pack).cuda()
        allp = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        moved = torch.tensor([ladder.rows_moved], dtype=torch.int64, device="cuda")
        dist.all_reduce(moved)
        ok = True
        if rank == 0:
            with Sampler(on.REGRESSION, (4, 5, 1), temps, S, si, device=local_rank, debug_traces=True, speculation=1, **kw) as one:
                one.set_data(tr, te)
                one.init_chains(w0)
                if mode == "replay":
                    one.replay(draws)
                else:
                    one.run()
                t1 = one.traces()
                sm1 = one.trace_summary(S // 2, S - S // 2)
                ns1, tot1, sw1 = one.swap_stats()
                st1 = one.get_state()
            ref = np.concatenate([t1["pos_w"].reshape(Rg, -1), t1["lik_prop"], t1["accept_list"], st1["w"], st1["eta"][:, None]], axis=1)
            got = torch.cat(allp).cpu().numpy()
            ok &= np.array_equal(got, ref)
            ok &= (ns, tot) == (ns1, tot1) and np.array_equal(sw, sw1)
            ok &= peer or int(moved.item()) > 0        # (the device-side exchange does not count rows on the host)
            ok &= bool(sw.any())
            for k in ("rmse_train", "rmse_test"):
                ok &= bool(np.allclose([pooled[k][q] for q in ("mean", "std", "min", "max")],
                                       [sm1[k][q] for q in ("mean", "std", "min", "max")], rtol=1e-9, atol=1e-12))
            ok &= bool(np.allclose(pooled["w_mean"], sm1["w_mean"], rtol=1e-9, atol=1e-12))
            ok &= bool(np.allclose(pooled["w_std"], sm1["w_std"], rtol=1e-7, atol=1e-10)) and pooled["n"] == sm1["n"]
            print("DIST_GPU_PASS %s mode=%s exchange=%s ok=%s world=%d swaps=%d/%d rows_moved=%d" % (tag, mode, "peer" if peer else "host", ok, world, ns, tot, int(moved.item())))
        return ok

    ok = one_pass(w0, "first")
    if os.environ.get("PT_TEST_REINIT", "0") == "1":
        # a second run on the SAME connected handles (ptfnn_init_chains again): the arrival flags still hold the first
        # run's counts and must not be taken for this run's
        ok = one_pass(np.random.RandomState(12).randn(Rg, cfg.P), "second") and ok
    if rank == 0:
        print("DIST_GPU_RESULT mode=%s exchange=%s ok=%s world=%d" % (mode, "peer" if peer else "host", ok, world))

    smp.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
