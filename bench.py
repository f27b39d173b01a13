#!/usr/bin/env python
"""bench.py -- replica MCMC steps/s of the Langevin parallel-tempering FNN sampler on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workloads (BASELINE.json configs):
  synth_ts  (default) synthetic 100k-point series, embed 4, FNN 4-64-1, 1024 temperatures per GPU
            (29 998 train / 19 998 test rows), Langevin l_prob 0.5, swap every 10 steps.  The ladder
            is partitioned over ranks: weak scaling, 1024 temperatures per GPU, boundary swaps over NVLink.
  sunspot   Sunspot 4-5-1, 10 temperatures, maxtemp 2, Langevin l_prob 0.5, swap every 50 steps (1 GPU)
  pendigit  PenDigit-shaped 16-256-10, 20k rows, 256 temperatures per GPU

One bench "step" = one swap interval of the whole ladder (swap_interval MCMC steps of every
replica + the swap round).  ``value`` = replica-steps/s with everything resident in HBM;
``e2e`` = the same through the public Sampler API with host buffers (dataset H2D + trace D2H every step).

The default line (workload synth_ts) also carries, under ``sub``, the other north-star configurations measured in
the same process -- ``burned_in`` (the same ladder after >= 2000 steps, where the acceptance rate is the
reference's steady state rather than ~1), ``strong_1024`` (the ladder of BASELINE configs[3] as worded: 1024
temperatures split over the N GPUs), and at N = 1 ``sunspot`` (the >= 100x target) and ``pendigit`` (the tcgen05
path) -- and, at N > 1, ``multi_gpu_parity``: a replay of the partitioned ladder compared bit for bit with a
single-GPU run of the whole ladder.  ``--no-sub`` skips them.
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_FLUSH_BYTES = 256 << 20
BURN_IN_STEPS = 2000


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
def load_workload(name: str, n_gpus: int, replicas_per_gpu=None):
    from ptnn_b200 import datasets
    from ptnn_b200.sampler import geometric_ladder
    if name == "synth_ts":
        train, test = datasets.synthetic_timeseries()
        per = replicas_per_gpu or 1024
        w = dict(task=0, topology=(4, 64, 1), swap_interval=10, learn_rate=0.01, l_prob=0.5, maxtemp=2,
                 desc="synthetic Mackey-Glass series 100k points, embed 4 lag 2, FNN 4-64-1, %d temperatures/GPU" % per)
    elif name == "sunspot":
        d = np.load(os.path.join(ROOT, "tests", "golden", "datasets.npz"))
        train, test = d["reg_Sunspot_train"], d["reg_Sunspot_test"]
        per = replicas_per_gpu or 10
        w = dict(task=0, topology=(4, 5, 1), swap_interval=50, learn_rate=0.1, l_prob=0.5, maxtemp=2,
                 desc="Sunspot one-step-ahead, FNN 4-5-1, %d temperatures, maxtemp 2 (BASELINE configs[0])" % per)
    elif name == "pendigit":
        train, test = datasets.synthetic_pendigit()
        per = replicas_per_gpu or 256
        w = dict(task=1, topology=(16, 256, 10), swap_interval=10, learn_rate=0.01, l_prob=0.5, maxtemp=10,
                 desc="PenDigit-shaped synthetic 16-256-10, 20k rows, %d temperatures/GPU" % per)
    else:
        raise SystemExit("unknown workload %s" % name)
    Rg = per * n_gpus
    w.update(name=name, train=train, test=test, R_per_gpu=per, R_global=Rg,
             temperatures=geometric_ladder(Rg, w["maxtemp"]) if Rg > 1 else np.ones(1))
    I, H, O = w["topology"]
    w["P"] = I * H + H * O + H + O
    return w


def config_of(w):
    """What both arms print as ``config``: the workload, nothing about how it was run."""
    return {"workload": w["desc"],
            "sampler": "Langevin PT, l_prob %.2f, lr %g, swap_interval %d, maxtemp %s" % (w["l_prob"], w["learn_rate"], w["swap_interval"], w["maxtemp"])}


def data_of(w):
    return "synthetic" if w["name"] != "sunspot" else "Sunspot (reference dataset), random-init weights"


def algorithmic_work(w):
    """logical passes of the reference per replica-step, fp32 storage."""
    I, H, O = w["topology"]
    N, M = w["train"].shape[0], w["test"].shape[0]
    row_bytes = 4 * (I + 1)
    fwd = 2 * (I * H + H * O) + (H + O)
    sgd = fwd + 4 * H * O + 2 * I * H + 5 * H + 5 * O
    sig = H + O
    trace = (w["P"] + 4) * 4
    # sfu = the algorithmic 2 transcendental ops per sigmoid (EX2 + RCP); sfu_issued = what the kernels issue: the
    # likelihood passes share one RCP among four hidden sigmoids (1.25 per sigmoid), the SGD recurrence does not
    lik_issued = H * 1.25 + O * 2
    rw = dict(flop=(N + M) * fwd, bytes=(N + M) * row_bytes + trace, sfu=(N + M) * sig * 2, sfu_issued=(N + M) * lik_issued,
              tensor_flop=(N + M) * 2 * (I * H + H * O))
    lg = dict(flop=2 * N * sgd + (N + M) * fwd, bytes=(3 * N + M) * row_bytes + trace, sfu=(3 * N + M) * sig * 2,
              sfu_issued=2 * N * sig * 2 + (N + M) * lik_issued, tensor_flop=(N + M) * 2 * (I * H + H * O))
    return rw, lg


# ------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        load = [v for v in sm if v > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the reference's multiprocessing path (oracle port)
# ------------------------------------------------------------------------------------------
def cpu_sample(w, n_procs, lx_pattern, seed=0):
    """A bounded sample of the workload on the host cores: ``n_procs`` replica processes (the coldest
    rungs of the ladder) run len(lx_pattern) steps each; lx_pattern forces Langevin (lx < l_prob) or
    random-walk steps so that the mix is exactly the workload's l_prob."""
    from oracle import cpu_baseline, ptfnn_numpy as on
    S = len(lx_pattern) + 1
    cfg = on.PTConfig(task=w["task"], topology=w["topology"], samples=S, swap_interval=w["swap_interval"],
                      use_langevin_gradients=True, l_prob=w["l_prob"], learn_rate=w["learn_rate"])
    R = n_procs
    draws = on.random_draws(cfg, R, seed, common_random_numbers=True)
    draws.lx[:] = np.asarray(lx_pattern)[None, :]
    w0 = np.random.RandomState(seed).randn(R, cfg.P)
    temps = np.resize(w["temperatures"], R)
    wall, units, _, _ = cpu_baseline.run(cfg, w["train"], w["test"], temps, w0, draws, S - 1)
    return wall, units


LG, RW = 0.0, 0.999


def cpu_sample_plan(w, cores):
    """-> (processes, lx pattern of one sample).  ~10-30 s of CPU work in total."""
    if w["name"] == "sunspot":                   # ~13 ms per replica-step: 10 procs x 120 steps ~ 16 s of CPU work
        return min(w["R_global"], 10), [LG, RW] * 60
    return cores, [LG, RW]                       # seconds per replica-step: one Langevin + one random-walk step per core


def cpu_baseline_record(w):
    from oracle import cpu_baseline
    cores = cpu_baseline.host_cores()
    n_procs, pattern = cpu_sample_plan(w, cores)
    wall, units = cpu_sample(w, n_procs, pattern)
    return {"value": units / wall, "unit": "replica-steps/s", "cores": cores, "kind": "port",
            "sample": "%d replica processes x %d steps (LG/RW alternating = l_prob 0.5) of the same workload, %.1f s wall" % (n_procs, len(pattern), wall)}


def run_reference_arm(args, w):
    """--impl reference: the reference's multiprocessing CPU path (oracle port; /root/reference does not
    exist on the GPU box) on the host cores, same workload / metric / unit.  Every bench step is a
    bounded sample; for the large workloads steps alternate all-Langevin / all-random-walk and the
    two mean times are combined with the workload's l_prob."""
    from oracle import cpu_baseline
    cores = cpu_baseline.host_cores()
    n_procs, pattern = cpu_sample_plan(w, cores)
    split = w["name"] != "sunspot"
    t_lg, t_rw, times, units = [], [], [], 0
    for k in range(args.warmup + args.steps):
        pat = ([LG] if k % 2 == 0 else [RW]) if split else pattern
        wall, u = cpu_sample(w, n_procs, pat, seed=k)
        if k >= args.warmup:
            times.append(wall); units += u
            (t_lg if k % 2 == 0 else t_rw).append(wall)
    if split and t_lg and t_rw:
        per_step = w["l_prob"] * statistics.mean(t_lg) + (1 - w["l_prob"]) * statistics.mean(t_rw)
        value = n_procs / per_step
        sample = "%d replica processes x 1 step per bench step, alternating all-Langevin / all-random-walk; value = procs / (l_prob*t_LG + (1-l_prob)*t_RW), t_LG %.2f s, t_RW %.2f s" % (
            n_procs, statistics.mean(t_lg), statistics.mean(t_rw))
    else:
        value = units / sum(times)
        sample = "%d replica processes x %d steps per bench step (LG/RW alternating = l_prob 0.5), swap rounds included" % (n_procs, len(pattern))
    line = {"impl": "reference", "metric": "replica_mcmc_steps_per_sec", "value": value, "unit": "replica-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": data_of(w), "config": config_of(w),
            "cpu_baseline": {"value": value, "unit": "replica-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "replica-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------
def measure_peaks(device):
    lib = ctypes.CDLL(os.path.join(ROOT, "parallel-tempering-neural-net_b200", "csrc", "libptfnn_peaks.so"))
    out = (ctypes.c_double * 4)()
    if lib.ptfnn_measure_peaks(int(device), out) != 0:
        return None
    return {"fp32_tflops": out[0], "mufu_gops": out[1], "smem_gbs": out[2], "fp32_3reg_tflops": out[3]}


def measured_peaks_file():
    mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp_path):
        d = json.load(open(mp_path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("bf16_tflops_sustained", 0) or 0)
    return 6650.0, "fallback 6.65 TB/s", 0.0


class Bench:
    """One process per GPU; every measurement of the B200 arm shares the device, the stream and the L2-flush buffer."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        from ptnn_b200 import capi
        if capi.device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: libptfnn has no CPU path")
        self.torch, self.args, self.rank, self.world, self.local_rank = torch, args, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=self.dev)
        self.peaks = None

    # ---- plumbing
    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def make(self, w, memo, samples, w0=None, state=None, distributed=True, **extra):
        """A sampler over this rank's block of ``w``'s ladder (the whole ladder when world == 1 or not distributed)."""
        from ptnn_b200.sampler import Sampler
        kw = dict(use_langevin_gradients=True, l_prob=w["l_prob"], learn_rate=w["learn_rate"], seed=self.args.seed,
                  common_random_numbers=True)
        if self.args.speculation:
            kw["speculation"] = self.args.speculation
        kw.update(extra)
        si = w["swap_interval"]
        if self.world > 1 and distributed:
            from ptnn_b200.distributed import make_gpu_ladder
            ladder, smp = make_gpu_ladder(w["task"], w["topology"], w["temperatures"], samples, si, device=self.local_rank,
                                          memoize_gradient=memo, **kw)
        else:
            smp = Sampler(w["task"], w["topology"], w["temperatures"], samples, si, device=self.local_rank,
                          memoize_gradient=memo, stream=self.stream, **kw)
            ladder = None
        smp.set_data(w["train"], w["test"])
        if w0 is None:
            w0 = np.random.RandomState(1000 + self.rank).randn(smp.R, smp.P)
        smp.init_chains(w0)
        if state is not None:                 # continue a burned-in chain: (w is w0) eta, lik, prior, tau
            smp.set_state(eta=state["eta"], lik=state["lik"], prior=state["prior"], tau=state["tau"])
        if ladder is not None:
            self.barrier()
        return smp, ladder

    def timed_steps(self, smp, ladder, si, n_warm, n_timed):
        """-> (sum of the timed steps' device time in ms, max over ranks; per-step ms of this rank)."""
        torch = self.torch
        adv = (lambda: ladder.run(si)) if ladder is not None else (lambda: smp.run(si))
        self.align(smp, ladder, si)
        for _ in range(n_warm):
            adv()
        self.barrier()
        self.first_timed = smp.step
        ev = []
        for _ in range(n_timed):
            self.flush.fill_(1)                                   # evict L2 between timed steps (untimed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            adv()
            b.record(self.stream)
            ev.append((a, b))
        self.barrier()
        ms = [a.elapsed_time(b) for a, b in ev]
        return self.max_over_ranks(sum(ms)), ms

    @staticmethod
    def align(smp, ladder, si):
        """Advance (untimed) so that every bench step ends ON a swap round: the regression chain swaps after steps
        i % si == 0 (R:427), so its bench steps are [k si + 1, (k+1) si]; the classification chain swaps after
        (i + 1) % si == 0 (C:438) and needs nothing.  A bench step that straddles a round would be two segments."""
        want = 1 if smp.cfg.task == 0 else 0
        off = (want - smp.step) % si
        if off:
            (ladder.run(off) if ladder is not None else smp.run(off))

    def acceptance(self, smp, first_step, n_steps):
        """Acceptance rate of steps [first_step, first_step + n_steps) of this rank's replicas: accept_list row i+1 is
        the number accepted BEFORE step i (R:380)."""
        a = smp.traces(first=first_step + 1, count=1, pos_w=False, debug=False)["accept_list"][:, 0]
        last = first_step + n_steps
        if smp.step > last:                                  # row last+1 (the count before step `last`) has been written
            b = smp.traces(first=last + 1, count=1, pos_w=False, debug=False)["accept_list"][:, 0]
        else:                                                # the chain stands right after step last-1
            b = smp.get_state()["num_accepted"].astype(np.float64)
        tot = self.sum_over_ranks(float(np.sum(b - a)))
        return tot / (n_steps * smp.R * self.world if self.dist is not None else n_steps * smp.R)

    def e2e(self, smp, ladder, w, n_warm, n_timed):
        """The same steps through the public API with HOST buffers: every step uploads the data set and reads the
        step's trace rows back.  The read-back is overlapped: step k's rows are fetched on the library's copy
        stream while step k+1 runs, and arrive as views (float32 weights, nothing widened)."""
        si = w["swap_interval"]
        adv = (lambda: ladder.run(si)) if ladder is not None else (lambda: smp.run(si))
        self.align(smp, ladder, si)
        for _ in range(n_warm):                                  # warm-up = the same calls (the first read-back page-locks its slots)
            smp.set_data(w["train"], w["test"])
            adv()
            smp.traces_end(smp.traces_begin(smp.step - si + 1, si, pos_w=True))
        self.barrier()
        t0 = time.perf_counter()
        d2h, pending, checksum = 0, None, 0.0
        for k in range(n_timed):
            smp.set_data(w["train"], w["test"])                  # H2D (host float64 arrays, as the reference passes them)
            adv()
            first = smp.step - si + 1
            ticket = smp.traces_begin(first, si, pos_w=True)     # D2H of this step, behind the kernel just launched
            if pending is not None:
                t = smp.traces_end(pending)                      # the previous step's rows: copied while this one runs
                checksum += float(t["lik_prop"][0, -1]) + float(t["pos_w"][0, -1, 0])
                d2h = sum(v.nbytes for v in t.values())
            pending = ticket
        t = smp.traces_end(pending)
        checksum += float(t["lik_prop"][0, -1]) + float(t["pos_w"][0, -1, 0])
        d2h = sum(v.nbytes for v in t.values())
        self.barrier()
        sec = self.max_over_ranks(time.perf_counter() - t0)
        I = w["topology"][0]
        ip = (I + 3) & ~3
        h2d = sum((n * ip + ((n + 3) & ~3)) * 4 for n in (w["train"].shape[0], w["test"].shape[0]))
        return {"value": w["R_global"] * si * n_timed / sec, "unit": "replica-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "readback": "overlapped (copy stream, page-locked double buffer, float32 views)",
                "finite": bool(np.isfinite(checksum))}

    def lx_mix(self, smp, w, first_step, n_steps):
        lx = smp.generate_draws(first_step, n_steps)[0][0]          # common random numbers: one lx per step
        n_lg = int(np.sum(lx < w["l_prob"]))
        return n_lg, n_steps - n_lg

    def rooflines(self, w, n_lg, n_rw, sec, K, clk):
        """Roofline records of chain_kernel for a timed region of n_lg Langevin + n_rw random-walk steps."""
        if self.peaks is None:
            self.peaks = measure_peaks(self.local_rank) or {}
        peaks, world = self.peaks, self.world
        rw, lg = algorithmic_work(w)
        Rg = w["R_global"]
        flops = Rg * (n_lg * lg["flop"] + n_rw * rw["flop"])
        byts = Rg * (n_lg * lg["bytes"] + n_rw * rw["bytes"])
        sfu = Rg * (n_lg * lg["sfu"] + n_rw * rw["sfu"])
        hbm_peak, hbm_src, bf16_peak = measured_peaks_file()
        fp32_peak = peaks.get("fp32_tflops")
        ach_tflops = flops / sec / 1e12 / world                 # per GPU
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(w["name"])
        kname = "chain_kernel<%d,%d,%d>" % tuple(w["topology"])
        fp32_roof = {"kernel": kname, "bound": "fp32-issue (not HBM: datasets are SMEM/L2 resident, SURVEY 8d)",
                     "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": (ach_tflops / fp32_peak) if fp32_peak else None,
                     "peak_source": "measured in this run (csrc/ptfnn_peaks.cu FMA microbenchmark)", "traffic": traffic,
                     "algorithmic_flop_per_launch": flops / K / world, "launch_ms": 1e3 * sec / K}
        mufu_peak = peaks.get("mufu_gops")
        ach_mufu = sfu / sec / 1e9 / world
        sfu_issued = Rg * (n_lg * lg["sfu_issued"] + n_rw * rw["sfu_issued"])
        mufu_roof = {"kernel": kname, "bound": "MUFU / XU pipe: 2 transcendental ops per sigmoid (not HBM: datasets are SMEM/L2 resident, SURVEY 8d)",
                     "achieved": ach_mufu / 1e3, "peak": (mufu_peak / 1e3) if mufu_peak else None, "unit": "Tops/s",
                     "frac": (ach_mufu / mufu_peak) if mufu_peak else None,
                     "peak_source": "measured in this run (csrc/ptfnn_peaks.cu MUFU microbenchmark: 16 lanes/clk/SM)", "traffic": traffic,
                     "algorithmic_sfu_ops_per_launch": sfu / K / world, "launch_ms": 1e3 * sec / K,
                     "issued_sfu_ops_per_launch": sfu_issued / K / world,
                     "frac_issued": (sfu_issued / sec / 1e9 / world / mufu_peak) if mufu_peak else None,
                     "note": "frac counts the algorithmic 2 ops per sigmoid; the likelihood passes issue 1.25 (one RCP per four sigmoids), frac_issued counts what is issued"}
        use_mufu = bool(mufu_roof["frac"] and fp32_roof["frac"] and mufu_roof["frac"] > fp32_roof["frac"])
        roofline = mufu_roof if use_mufu else fp32_roof
        alt = {
            "hbm": {"bound": "hbm", "achieved": byts / sec / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
                    "frac": byts / sec / 1e9 / world / hbm_peak, "peak_source": hbm_src,
                    "algorithmic_bytes_per_launch": byts / K / world},
            ("fp32_issue" if use_mufu else "mufu"): (fp32_roof if use_mufu else mufu_roof),
            "smem_broadcast": {"achieved_gbs": byts / sec / 1e9 / world, "peak_gbs": peaks.get("smem_gbs")},
        }
        # The Langevin steps are a serial recurrence over the training rows: their bound is the
        # dependent-issue latency of one row, not a throughput peak.  floor = sum of the measured latencies of the
        # instructions on the chain (tools/latency_probe.cu); the upper bound charges the WHOLE
        # timed region (random-walk steps included) to the serial rows.
        chain_floor = {(4, 64, 1): 176, (4, 5, 1): 138, (4, 10, 1): 150, (16, 256, 10): 340}.get(tuple(w["topology"]))
        serial_rows = n_lg * 2 * w["train"].shape[0]
        # (only where one CTA per temperature runs the steps one after the other: with speculative windows -- small
        # ladders, or a ladder that leaves CTA slots free -- several steps of a chain are evaluated at once)
        sequential = w["name"] == "pendigit" or w["R_per_gpu"] >= 1024
        if sequential and chain_floor and serial_rows and clk and clk.get("sm_mhz"):
            cyc = sec / serial_rows * clk["sm_mhz"] * 1e6
            alt["serial_chain"] = {"bound": "dependent-issue latency of the SGD recurrence (one chain per temperature)",
                                   "serial_rows_per_temperature": serial_rows, "cycles_per_row_upper_bound": cyc,
                                   "floor_cycles_per_row": chain_floor, "frac": chain_floor / cyc}
        if tuple(w["topology"]) == (16, 256, 10) and bf16_peak:
            # K5: both layers of the likelihood pass on the tf32 tensor pipe (3xTF32); dense tf32 peak = half the bf16 one
            tf = Rg * (n_lg * lg["tensor_flop"] + n_rw * rw["tensor_flop"]) / sec / 1e12 / world
            alt["tensor_tf32"] = {"bound": "tensor", "achieved": tf, "peak": bf16_peak / 2, "unit": "TFLOP/s", "frac": tf / (bf16_peak / 2),
                                  "peak_source": "MEASURED_PEAKS.json bf16 sustained / 2",
                                  "note": "algorithmic flop of the two dense layers (one logical pass); the kernel issues 3 tf32 products per logical one"}
        return roofline, alt

    # ---- one workload, measured like the headline: value, memoised value, e2e, acceptance, roofline
    def measure(self, w, K, W, with_cpu, with_pipeline=False, distributed=True, burn=None):
        """``burn``: dict(w=..., eta=..., lik=..., prior=..., tau=...) of a burned-in ladder (this rank's block) to continue."""
        si = w["swap_interval"]
        S = si * (K + W + 1) + 2
        w0 = None if burn is None else burn["w"]
        clocks = ClockSampler(self.local_rank) if self.rank == 0 else None
        smp, ladder = self.make(w, 0, S, w0=w0, state=burn, distributed=distributed)
        total_ms, ms_list = self.timed_steps(smp, ladder, si, W, K)
        clk = clocks.stop() if clocks else None
        world = self.world if distributed else 1
        units = w["R_global"] * si
        value = units * K / (total_ms * 1e-3)
        n_lg, n_rw = self.lx_mix(smp, w, self.first_timed, K * si)
        acc = self.acceptance(smp, self.first_timed, K * si) if distributed or self.rank == 0 else None
        smp.close()
        smp, ladder = self.make(w, 1, S, w0=w0, state=burn, distributed=distributed)
        total_ms_memo, _ = self.timed_steps(smp, ladder, si, W, K)
        smp.close()
        smp, ladder = self.make(w, 0, S, w0=w0, state=burn, distributed=distributed)
        e2e = self.e2e(smp, ladder, w, W, K)
        pipe = None
        if with_pipeline and self.rank == 0:
            n_rows = smp.step
            smp.trace_summary(1, n_rows)
            got = []
            for _ in range(3):
                self.flush.fill_(1)
                self.torch.cuda.synchronize(self.dev)
                got.append(smp.trace_summary(1, n_rows))
            ms = float(np.mean([g["kernel_ms"] for g in got]))
            hbm_peak, hbm_src, _ = measured_peaks_file()
            ach = got[0]["bytes_read"] / (ms * 1e-3) / 1e9
            pipe = {"kernel": "trace_summary_kernel<2>", "bound": "hbm", "rows_pooled": int(got[0]["n"]),
                    "algorithmic_bytes_per_launch": int(got[0]["bytes_read"]), "launch_ms": ms, "achieved": ach, "unit": "GB/s",
                    "peak": hbm_peak, "frac": ach / hbm_peak, "peak_source": hbm_src}
        smp.close()
        if self.rank != 0:
            return None
        roofline, alt = self.rooflines(w, n_lg, n_rw, total_ms * 1e-3, K, clk)
        rec = {"value": value, "unit": "replica-steps/s", "ms_per_step": total_ms / K, "steps": K, "warmup": W,
               "config": config_of(w), "data": data_of(w), "n_gpus": world,
               "timed_region": {"replicas_total": w["R_global"], "replica_steps_per_bench_step": units,
                                "langevin_steps": n_lg, "random_walk_steps": n_rw, "acceptance_rate": acc,
                                "memoize_gradient": 0, "l2": "flushed between timed steps (256 MiB write)",
                                "first_timed_chain_step": self.first_timed, "bench_step": "swap_interval chain steps ending on a swap round",
                                "start": "random initial weights" if burn is None else "continued after %d burn-in steps" % BURN_IN_STEPS},
               "value_memoized": units * K / (total_ms_memo * 1e-3), "ms_per_step_memoized": total_ms_memo / K,
               "e2e": e2e, "gpu_launches": K, "clocks": clk, "roofline": roofline, "roofline_alt": alt, "step_ms": ms_list}
        if with_cpu:
            rec["cpu_baseline"] = cpu_baseline_record(w)
        if pipe:
            rec["result_pipeline"] = pipe
        return rec

    # ---- burn-in: >= BURN_IN_STEPS steps of the same ladder (memoised gradient: same chain, fewer passes), state kept on the host
    def burn_in(self, w):
        S = BURN_IN_STEPS + 2
        smp, ladder = self.make(w, 1, S)
        t0 = time.perf_counter()
        (ladder.run(BURN_IN_STEPS) if ladder is not None else smp.run(BURN_IN_STEPS))
        st = smp.get_state()
        sec = time.perf_counter() - t0
        a = smp.traces(first=S - 200, count=1, pos_w=False, debug=False)["accept_list"][:, 0]      # accepted before step S-201
        late = self.sum_over_ranks(float(np.sum(st["num_accepted"] - a))) / (200.0 * smp.R * (self.world if self.dist is not None else 1))
        ns, tot, _ = smp.swap_stats(0)
        smp.close()
        st["burn_seconds"], st["acceptance_last_200"], st["swap_rate"] = sec, late, (ns / tot if tot else None)
        return st

    # ---- N > 1: the partitioned ladder against ONE GPU holding the whole ladder, bit for bit
    def multi_gpu_parity(self, w, n_steps=21):
        """Replay of ``n_steps`` steps (two swap rounds) with the bench's own 1024 temperatures per rank: the draws are the
        Philox draws dumped by the library, the swap uniforms are scaled down so that most pairs swap and vectors
        cross every rank boundary (R:741-748, R:674).  Every rank hashes its block of the result; rank 0 runs the whole
        ladder on its GPU alone and compares the hashes of the corresponding blocks."""
        from types import SimpleNamespace
        dist, world, rank = self.dist, self.world, self.rank
        si, Rg = w["swap_interval"], w["R_global"]
        S = n_steps + 1
        rounds = sum(1 for i in range(n_steps) if i % si == 0 and i != 0) + 1
        u_swap = (np.random.RandomState(77).rand(rounds, Rg - 1) * 0.05).astype(np.float32)
        w0 = np.random.RandomState(4242).randn(Rg, w["P"])
        per = Rg // world

        def digest(smp, lo, n):
            t = smp.traces(pos_w=True, debug=False)
            st = smp.get_state()
            h = []
            for k in range(0, n, per):
                m = hashlib.sha256()
                for a in (t["pos_w"][k:k + per], t["lik_prop"][k:k + per], t["accept_list"][k:k + per], st["w"][k:k + per], st["eta"][k:k + per]):
                    m.update(np.ascontiguousarray(a).tobytes())
                h.append(m.hexdigest())
            return h

        smp, ladder = self.make(w, 0, S, w0=w0[rank * per:(rank + 1) * per])
        lx, z, z_eta, u = smp.generate_draws(0, n_steps)
        ladder.run(None, SimpleNamespace(lx=lx, z=z, z_eta=z_eta, u=u), u_swap)
        mine = digest(smp, rank * per, per)
        ns, tot, sw = smp.swap_stats()
        smp.close()
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        out = None
        if rank == 0:
            one, _ = self.make(w, 0, S, w0=w0, distributed=False)
            lx, z, z_eta, u = one.generate_draws(0, n_steps)
            one.replay(SimpleNamespace(lx=lx, z=z, z_eta=z_eta, u=u, u_swap=u_swap))
            ref = digest(one, 0, Rg)
            ns1, tot1, sw1 = one.swap_stats()
            one.close()
            got = [h for hs in allh for h in hs]
            crossings = [int(sw[:, g * per - 1].sum()) for g in range(1, world)]
            out = {"ok": bool(got == ref and (ns, tot) == (ns1, tot1) and np.array_equal(sw, sw1)),
                   "temperatures_per_rank": per, "ranks": world, "steps": n_steps, "swap_rounds": int(sw.shape[0]),
                   "swaps": int(ns), "swap_proposals": int(tot), "swaps_across_each_rank_boundary": crossings,
                   "compared": "sha256 of pos_w, lik_prop, accept_list, final w and eta of every rank's block vs the same block of a single-GPU run of all %d temperatures" % Rg}
        self.barrier()
        return out


def run_b200_arm(args, w, rank, world, local_rank):
    b = Bench(args, rank, world, local_rank)
    K, W = args.steps, args.warmup
    head = b.measure(w, K, W, with_cpu=(world == 1 and not args.no_cpu_baseline), with_pipeline=True)
    sub = {}
    parity = None
    if args.sub and args.workload == "synth_ts":
        # ---- the same ladder in the regime the reference runs in (acceptance 12-30 %)
        st = b.burn_in(w)
        rec = b.measure(w, K, W, with_cpu=False, burn=st)
        if rec is not None:
            rec["burn_in"] = {"steps": BURN_IN_STEPS, "seconds": st["burn_seconds"], "acceptance_rate_last_200_steps": st["acceptance_last_200"],
                              "swap_rate": st["swap_rate"]}
            sub["burned_in"] = rec
        # ---- BASELINE configs[3] as worded: ONE ladder of 1024 temperatures split over the N GPUs (strong scaling)
        if world > 1 and 1024 % world == 0:
            ws = load_workload("synth_ts", world, 1024 // world)
            rec = b.measure(ws, K, W, with_cpu=False)
            if rec is not None:
                rec["scaling"] = "strong"
                sub["strong_1024"] = rec
            st = b.burn_in(ws)                       # ... and in the reference's steady state (speculative windows follow the acceptance rate)
            rec = b.measure(ws, K, W, with_cpu=False, burn=st)
            if rec is not None:
                rec["scaling"] = "strong"
                rec["burn_in"] = {"steps": BURN_IN_STEPS, "seconds": st["burn_seconds"], "acceptance_rate_last_200_steps": st["acceptance_last_200"],
                                  "swap_rate": st["swap_rate"]}
                sub["strong_1024_burned_in"] = rec
            parity = b.multi_gpu_parity(w)
        elif world == 1:
            sub["strong_1024"] = {"same_as": "the headline line: at one GPU the 1024-temperature ladder of configs[3] is the weak-scaling workload"}
        # ---- the reference's own run (the >= 100x target) and the tcgen05 configuration: single-GPU workloads
        if world == 1:
            ws = load_workload("sunspot", 1)
            sub["sunspot"] = b.measure(ws, 80, 20, with_cpu=not args.no_cpu_baseline)
            wp = load_workload("pendigit", 1)
            sub["pendigit"] = b.measure(wp, max(3, min(K, 10)), 3, with_cpu=not args.no_cpu_baseline)
    if rank == 0:
        line = {"metric": "replica_mcmc_steps_per_sec", "value": head["value"], "unit": "replica-steps/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": head["data"], "config": head["config"],
                "timed_region": head["timed_region"],
                "parallelism": "ladder partitioned over %d GPU(s), swap rounds in-kernel over NVLink peer memory" % world if world > 1 else "1 GPU, in-kernel swap round",
                "value_memoized": head["value_memoized"], "ms_per_step_memoized": head["ms_per_step_memoized"],
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "roofline": head["roofline"],
                "roofline_alt": head["roofline_alt"], "cpu_baseline": head.get("cpu_baseline"), "peaks_measured": b.peaks,
                "step_ms": head["step_ms"]}
        if "result_pipeline" in head:
            line["result_pipeline"] = head["result_pipeline"]
        if sub:
            line["sub"] = sub
        if parity is not None:
            line["multi_gpu_parity"] = parity["ok"]
            line["multi_gpu_parity_detail"] = parity
        print(json.dumps(line))
    if b.dist is not None:
        b.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synth_ts", choices=["synth_ts", "sunspot", "pendigit"])
    ap.add_argument("--replicas-per-gpu", type=int, default=None)
    ap.add_argument("--ladder-total", type=int, default=None,
                    help="strong scaling: a fixed ladder of this many temperatures split over the GPUs "
                         "(BASELINE configs[3] as worded: 1024 over 1/2/4/8); default is weak scaling, 1024 per GPU")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--speculation", type=int, default=0, help="CTAs per temperature: 0 = the library's automatic choice (default), 1 = none, K = fixed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", dest="sub", action="store_false", help="only the headline measurement (no sub-records, no parity replay)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    n_dev = max(world, 1) if args.impl == "b200" else args.gpus
    args.scaling = "weak"
    if args.ladder_total:
        if args.ladder_total % n_dev:
            raise SystemExit("--ladder-total %d does not split evenly over %d GPUs" % (args.ladder_total, n_dev))
        args.replicas_per_gpu, args.scaling = args.ladder_total // n_dev, "strong"
    if args.replicas_per_gpu or args.ladder_total:
        args.sub = False
    w = load_workload(args.workload, n_dev, args.replicas_per_gpu)
    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, w)
        return
    run_b200_arm(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
