"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the reference's multiprocessing CPU path, restated.

Used by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm on the GPU box, where
/root/reference does not exist.  Structure follows the reference: one OS process per replica
(``ptReplica(multiprocessing.Process)``, R:138) running the float64 row-by-row NumPy chain of
oracle/ptfnn_numpy.py, and a coordinator that, every ``swap_interval`` steps, collects
[w, eta, lhood] from every replica, runs the sequential sweep and hands the vectors back
(R:427-437, R:719-752) -- pipes instead of Queue/Event pairs.  ``kind`` = "port".
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from . import ptfnn_numpy as on


def _worker(conn, cfg, train, test, temperature, w0, lx, z, z_eta, u):
    rep = on.Replica(cfg, train, test, temperature, w0)
    S = cfg.samples
    tr = on.new_traces(1, S, cfg.P, 1)
    conn.send(("ready", None))
    i = 0
    while True:
        msg, arg = conn.recv()
        if msg == "run":                       # run `arg` steps, then report the hand-shake vector
            t0 = time.perf_counter()
            for _ in range(arg):
                rep.step(i, lx[i], z[i], z_eta[i], u[i], tr, 0)
                i += 1
            conn.send((rep.w, rep.eta, rep.swap_field(), time.perf_counter() - t0))
        elif msg == "set":                     # R:435-437: only w and eta are taken back
            rep.w, rep.eta = arg
        elif msg == "stop":
            conn.send((tr.accept_list[0, i], float(tr.rmse_train[0, i]), float(tr.rmse_test[0, i])))
            conn.close()
            return


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run(cfg: on.PTConfig, train, test, temperatures, w0, draws: on.Draws, n_steps: int, max_procs=None):
    """Run ``n_steps`` steps of every replica, one process per replica (all started at once, as the
    reference does), with swap rounds where due.  Returns (wall seconds of the sampling phase,
    replica-steps executed, swaps, per-replica summaries)."""
    R = len(temperatures)
    ctx = mp.get_context("fork")
    conns, procs = [], []
    for k in range(R):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(b, cfg, train, test, temperatures[k], w0[k], draws.lx[k], draws.z[k],
                                              draws.z_eta[k], draws.u[k]), daemon=True)
        p.start()
        conns.append(a)
        procs.append(p)
    for c in conns:
        c.recv()                               # init (R:266-285) is outside the timed sampling phase
    t0 = time.perf_counter()
    i, rnd, num_swap = 0, 0, 0
    while i < n_steps:
        seg = 0                                # steps up to and including the next swap step
        while i + seg < n_steps:
            seg += 1
            if cfg.swap_due(i + seg - 1):
                break
        for c in conns:
            c.send(("run", seg))
        vecs = [c.recv() for c in conns]
        i += seg
        if cfg.swap_due(i - 1) and R > 1:
            src, sw = on.swap_sweep([v[2] for v in vecs], draws.u_swap[rnd % len(draws.u_swap)])
            for k, c in enumerate(conns):
                c.send(("set", (vecs[src[k]][0], vecs[src[k]][1])))
            num_swap += sum(sw)
            rnd += 1
    wall = time.perf_counter() - t0
    out = []
    for c in conns:
        c.send(("stop", None))
        out.append(c.recv())
    for p in procs:
        p.join()
    return wall, R * n_steps, num_swap, out
