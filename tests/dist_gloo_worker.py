"""Worker for tests/test_distributed_gloo.py (launched under torch.distributed.run, gloo, CPU).

Exercises the PRODUCT's ladder-partition / boundary-exchange logic (ptnn_b200.distributed.
PartitionedLadder: all_gather of the swap fields, identical sequential sweep on every rank,
isend/irecv of only the rows that cross a rank boundary) with oracle-backed chains standing in for
the GPU block, and checks on rank 0 that the sharded run is bit-identical to the single-process
oracle run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ptfnn_numpy as on                      # noqa: E402
from ptnn_b200.distributed import PartitionedLadder, partition   # noqa: E402
from tests import common as cm                            # noqa: E402


class OracleChains:
    """Same interface as ptnn_b200.distributed.GpuChains, backed by oracle replicas (CPU, float64)."""

    def __init__(self, cfg, train, test, temps, w0, lo, n, Rg):
        self.cfg, self.lo, self.R, self.Rg = cfg, lo, n, Rg
        self.reps = [on.Replica(cfg, train, test, temps[lo + k], w0[lo + k]) for k in range(n)]
        self.tr = on.new_traces(n, cfg.samples, cfg.P, 1)
        self.step, self.last_step = 0, cfg.samples - 1
        self.rounds_done = 0
        self._pending, self._final = False, False
        self.lhood_global = torch.zeros(Rg, dtype=torch.float64)
        self.rows_in = torch.zeros(n, cfg.P + 1, dtype=torch.float64)
        self.swaps = []

    def run(self, n, draws):
        done = 0
        while done < n and self.step < self.last_step:
            i = self.step
            for k, rep in enumerate(self.reps):
                rep.step(i, draws.lx[k, i], draws.z[k, i], draws.z_eta[k, i], draws.u[k, i], self.tr, k)
            self.step += 1
            done += 1
            if self.cfg.swap_due(i):
                self._pending, self._final = True, False
                return done
        if self.step == self.last_step and self.rounds_done < self.cfg.total_rounds():
            self._pending, self._final = True, True
        return done

    def swap_pending(self):
        return self._pending, self._final

    def swap_export(self):
        if self._final:
            lh = [rep.likelihood for rep in self.reps]
        else:
            lh = [rep.swap_field() for rep in self.reps]
        rows = np.stack([np.concatenate([rep.w, [rep.eta]]) for rep in self.reps])
        return torch.tensor(lh, dtype=torch.float64), torch.from_numpy(rows.copy())

    def swap_plan(self, lhood_global, u_row):
        src, sw = on.swap_sweep(lhood_global.numpy().tolist(), u_row)
        self.swaps.append(sw)
        self.rounds_done += 1
        if self._final:
            self._pending = False
        return np.asarray(src)

    def swap_apply(self, src, rows_local, rows_in):
        new = []
        for k in range(self.R):
            s = int(src[self.lo + k])
            row = rows_local[s - self.lo] if self.lo <= s < self.lo + self.R else rows_in[k]
            new.append(row.numpy().copy())
        for k, rep in enumerate(self.reps):
            rep.w, rep.eta = new[k][:-1].copy(), float(new[k][-1])
        self._pending = False
        if self.step == self.last_step and self.rounds_done < self.cfg.total_rounds():
            self._pending, self._final = True, True


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    Rg, S = int(os.environ.get("PT_TEST_R", "4")), int(os.environ.get("PT_TEST_S", "31"))
    cfg = on.PTConfig(task=on.REGRESSION, topology=(4, 5, 1), samples=S, swap_interval=5,
                      use_langevin_gradients=True, l_prob=0.5, learn_rate=0.1)
    temps = on.geometric_ladder(Rg, 2)
    w0 = np.random.RandomState(11).randn(Rg, cfg.P)
    draws = on.random_draws(cfg, Rg, seed=5, common_random_numbers=False)
    # make swaps frequent so rows really cross the rank boundary
    draws.u_swap[:] = draws.u_swap * 0.3
    lo, n = partition(Rg, world, rank)
    local = on.Draws(lx=draws.lx[lo:lo + n], z=draws.z[lo:lo + n], z_eta=draws.z_eta[lo:lo + n],
                     u=draws.u[lo:lo + n], u_swap=None)
    chains = OracleChains(cfg, tr, te, temps, w0, lo, n, Rg)
    ladder = PartitionedLadder(chains, Rg, lo, n)
    ladder.run(None, local, draws.u_swap)
    # gather final (w, eta) and accept counts on rank 0
    mine = torch.from_numpy(np.stack([np.concatenate([rep.w, [rep.eta, rep.num_accepted]]) for rep in chains.reps]))
    allrows = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allrows, mine)
    moved = torch.tensor([ladder.rows_moved], dtype=torch.int64)
    dist.all_reduce(moved)
    ok = True
    if rank == 0:
        ref = on.run_pt(cfg, tr, te, temps, w0, draws)
        got = torch.cat(allrows).numpy()
        ok &= np.array_equal(got[:, :-2], ref.final_w)
        ok &= np.array_equal(got[:, -2], ref.final_eta)
        ok &= np.array_equal(got[:, -1], ref.accept_list[:, -1] + ref.accepted[:, -1])
        ok &= np.array_equal(np.asarray(chains.swaps, dtype=bool), ref.swapped)
        ok &= ladder.rounds == cfg.total_rounds()
        ok &= int(moved.item()) > 0                        # the boundary really was crossed
        print("DIST_GLOO_RESULT ok=%s rounds=%d rows_moved=%d num_swap=%d" % (ok, ladder.rounds, int(moved.item()), ref.num_swap))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
