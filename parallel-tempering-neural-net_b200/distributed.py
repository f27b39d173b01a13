"""Ladder partitioned over ranks (one process per GPU, torch.distributed).

SURVEY 8(e): replicas are independent between swap rounds, so rank g holds a contiguous block of
R/G temperatures (ladder order = rank order) and the only communication is the swap round:

  1. every rank contributes the swap field of its replicas (R:430 / C:439)  -> all_gather (R floats)
  2. every rank runs the reference's SEQUENTIAL sweep (R:741-748) on the gathered vector with the
     same uniforms (replay) or the same Philox counters (free-running) -> identical ``src``
  3. only rows whose source slot lives on another rank move: batched isend/irecv of (w, eta),
     (P+2) float32 words each (eta as its float64 bit pattern) -- in expectation the rows next to the G-1 rank boundaries
  4. install, continue.

On GPUs that can map each other's memory (NVLink / NVSwitch, one box) steps 1-4 run INSIDE the
persistent kernel (``peer=True``, the default of ``make_gpu_ladder``): the ranks exchange CUDA IPC
handles of their swap windows once, then every round is a push of R/G lhood fields to the peers, a
flag per rank, the sweep on local memory and a pull of the few rows that cross a rank boundary --
no host round trip and no collective call (csrc/ptfnn_kernels.cuh: peer_exchange_lhood).  The
host-completed round below stays as the portable path.

``PartitionedLadder`` holds that logic over an abstract ``chains`` object so that it is exercised
on CPU with the gloo backend (tests/test_distributed_gloo.py, oracle-backed chains) and on GPUs
with NCCL (``GpuChains`` over libptfnn).  The data path has no other collective.
"""
from __future__ import annotations

import numpy as np


def partition(n_global: int, world: int, rank: int):
    """Contiguous equal blocks (all_gather needs equal counts)."""
    if n_global % world != 0:
        raise ValueError("ladder of %d temperatures does not split evenly over %d ranks" % (n_global, world))
    per = n_global // world
    return rank * per, per


SERIES = ("rmse_train", "rmse_test", "acc_train", "acc_test")


def combine_summaries(parts):
    """Pool the per-rank results of ``Sampler.trace_summary`` (each over that rank's block of the ladder)
    into the statistics of the whole ladder: what the reference computes over all chains at once
    (R:1036-1044).  Exact up to fp64 rounding: mean = sum n_k m_k / n, var = sum n_k (s_k^2 + (m_k - mean)^2) / n."""
    parts = [p for p in parts if p is not None and p["n"] > 0]
    n = float(sum(p["n"] for p in parts))
    out = {"n": int(n)}

    def pool(ms, ss):
        ms, ss = np.asarray(ms, dtype=np.float64), np.asarray(ss, dtype=np.float64)
        wts = np.asarray([p["n"] for p in parts], dtype=np.float64).reshape((-1,) + (1,) * (ms.ndim - 1)) / n
        mean = (wts * ms).sum(axis=0)
        var = (wts * (ss * ss + (ms - mean) ** 2)).sum(axis=0)
        return mean, np.sqrt(var)

    for k in SERIES:
        mean, std = pool([p[k]["mean"] for p in parts], [p[k]["std"] for p in parts])
        out[k] = {"mean": float(mean), "std": float(std), "min": min(p[k]["min"] for p in parts),
                  "max": max(p[k]["max"] for p in parts)}
    if all(p.get("w_mean") is not None for p in parts):
        out["w_mean"], out["w_std"] = pool([p["w_mean"] for p in parts], [p["w_std"] for p in parts])
    else:
        out["w_mean"] = out["w_std"] = None
    return out


class GpuChains:
    """The local block of the ladder on this rank's GPU (libptfnn handle + torch swap buffers)."""

    def __init__(self, sampler, torch_device):
        import torch
        self.torch = torch
        self.s = sampler
        self.device = torch_device
        R, Rg, P = sampler.R, sampler.Rg, sampler.P
        self.lhood_local = torch.zeros(R, dtype=torch.float64, device=torch_device)
        self.lhood_global = torch.zeros(Rg, dtype=torch.float64, device=torch_device)
        self.rows_local = torch.zeros(R, P + 2, dtype=torch.float32, device=torch_device)
        self.rows_in = torch.zeros(R, P + 2, dtype=torch.float32, device=torch_device)
        sampler.set_stream(torch.cuda.current_stream(torch_device).cuda_stream)

    @property
    def step(self):
        return self.s.step

    @property
    def last_step(self):
        return self.s.S - 1

    def run(self, n, draws=None):
        return self.s.replay(draws, n_steps=n) if draws is not None else self.s.run(n)

    def swap_pending(self):
        return self.s.swap_pending()

    def swap_export(self):
        self.s.swap_export(self.lhood_local.data_ptr(), self.rows_local.data_ptr())
        return self.lhood_local, self.rows_local

    def swap_plan(self, lhood_global, u_row):
        return self.s.swap_plan(lhood_global.data_ptr(), u_row)[0]

    def swap_apply(self, src, rows_local, rows_in):
        self.s.swap_apply(src, rows_local.data_ptr(), rows_in.data_ptr())


class PartitionedLadder:
    def __init__(self, chains, n_global: int, offset: int, n_local: int, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.chains = chains
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.Rg, self.lo, self.R = n_global, offset, n_local
        self.per = n_global // self.world
        self.rows_moved = 0          # (w, eta) rows this rank received from other ranks
        self.rounds = 0

    def owner(self, slot: int) -> int:
        return slot // self.per

    def swap_round(self, u_row=None, final=False):
        dist = self.dist
        lh_local, rows_local = self.chains.swap_export()
        lh_global = self.chains.lhood_global
        dist.all_gather_into_tensor(lh_global, lh_local, group=self.group)
        src = self.chains.swap_plan(lh_global, u_row)
        self.rounds += 1
        if final:
            return src
        ops = []
        rows_in = self.chains.rows_in
        for g in range(self.Rg):                      # ascending g on every rank: pairwise order matches
            s = int(src[g])
            dst_rank, src_rank = self.owner(g), self.owner(s)
            if dst_rank == src_rank:
                continue
            if dst_rank == self.rank:
                ops.append(dist.P2POp(dist.irecv, rows_in[g - self.lo], src_rank, group=self.group, tag=g))
                self.rows_moved += 1
            elif src_rank == self.rank:
                ops.append(dist.P2POp(dist.isend, rows_local[s - self.lo], dst_rank, group=self.group, tag=g))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.chains.swap_apply(src, rows_local, rows_in)
        return src

    def init_chains(self, w0_local):
        """(Re-)initialise this rank's block.  On peer-connected handles the ranks meet at a barrier afterwards:
        no rank may publish into the swap window of a rank that is still reading the previous run's
        (include/ptfnn.h, ptfnn_peer_connect)."""
        self.chains.s.init_chains(w0_local)
        self.rounds = 0
        self.rows_moved = 0
        self.chains.s.sync()
        self.dist.barrier(self.group)

    def summary(self, first=0, count=None, posterior=True):
        """Result statistics of the WHOLE ladder: every rank reduces its own traces on its
        GPU (Sampler.trace_summary), the few numbers per rank are all-gathered and pooled; identical on
        every rank."""
        mine = self.chains.s.trace_summary(first, count, posterior)
        mine = {k: v for k, v in mine.items() if k not in ("kernel_ms", "bytes_read")}
        parts = [None] * self.world
        self.dist.all_gather_object(parts, mine, group=self.group)
        return combine_summaries(parts)

    def run(self, n_steps=None, draws=None, u_swap=None):
        """Advance every rank's block by up to ``n_steps`` steps, completing the swap rounds that fall
        due.  ``u_swap`` (replay): [rounds, Rg-1] indexed by absolute round number."""
        ch = self.chains
        if getattr(ch, "peer", False):                   # rounds complete on the device
            todo = ch.last_step - ch.step if n_steps is None else min(n_steps, ch.last_step - ch.step)
            if draws is None:
                return ch.s.run(todo)
            from types import SimpleNamespace
            return ch.s.replay(SimpleNamespace(lx=draws.lx, z=draws.z, z_eta=draws.z_eta, u=draws.u, u_swap=u_swap),
                               n_steps=todo)
        todo = ch.last_step - ch.step if n_steps is None else min(n_steps, ch.last_step - ch.step)
        done = 0
        while done < todo:
            k = ch.run(todo - done, draws)
            if k <= 0:
                break
            done += k
            while True:                                  # a chain ending on a swap step owes two rounds
                pending, final = ch.swap_pending()
                if not pending:
                    break
                u = None if u_swap is None else np.asarray(u_swap)[self.rounds]
                self.swap_round(u, final)
        return done


def make_gpu_ladder(task, topology, temperatures_global, samples, swap_interval, *, group=None, device=None,
                    peer=True, **sampler_kw):
    """One call per rank: builds the local Sampler for this rank's block and the exchange logic.
    ``peer=True``: swap rounds complete on the device through peer memory (one box, NVLink);
    ``peer=False``: the host completes them with all_gather + isend/irecv."""
    import torch
    import torch.distributed as dist
    from . import capi
    from .sampler import Sampler
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    temps = np.asarray(temperatures_global, dtype=np.float64)
    lo, n = partition(len(temps), world, rank)
    dev_index = torch.cuda.current_device() if device is None else device
    if world > 1 and not capi.has_topology(task, topology):
        # a topology that is compiled on demand: rank 0 compiles it once, the others load the cached library afterwards
        if rank == 0:
            capi.ensure_topology(task, topology)
        dist.barrier(group)
    smp = Sampler(task, topology, temps[lo:lo + n], samples, swap_interval, n_replicas_global=len(temps),
                  replica_offset=lo, device=dev_index, **sampler_kw)
    chains = GpuChains(smp, torch.device("cuda", dev_index))
    chains.peer = False
    if peer and world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, smp.peer_export(), group=group)
        smp.peer_connect(handles, rank)
        chains.peer = True
        dist.barrier(group)                              # every rank has mapped every window
    return PartitionedLadder(chains, len(temps), lo, n, group), smp
