#!/usr/bin/env python
"""bench.py -- replica MCMC steps/s of the Langevin parallel-tempering FNN sampler on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workloads (BASELINE.json configs):
  synth_ts  (default) synthetic 100k-point series, embed 4, FNN 4-64-1, 1024 temperatures per GPU
            (29 998 train / 19 998 test rows), Langevin l_prob 0.5, swap every 10 steps.  The ladder
            is partitioned over ranks: weak scaling, 1024 temperatures per GPU, boundary swaps over NCCL.
  sunspot   Sunspot 4-5-1, 10 temperatures, maxtemp 2, Langevin l_prob 0.5, swap every 50 steps (1 GPU)
  pendigit  PenDigit-shaped 16-256-10, 20k rows, 256 temperatures per GPU

One bench "step" = one swap interval of the whole ladder (swap_interval MCMC steps of every
replica + the swap round).  ``value`` = replica-steps/s with everything resident in HBM;
``e2e`` = the same through the public Sampler API with host buffers (dataset H2D + trace D2H every step).
"""
from __future__ import annotations

import argparse
import ctypes
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


def algorithmic_work(w):
    """logical passes of the reference per replica-step, fp32 storage."""
    I, H, O = w["topology"]
    N, M = w["train"].shape[0], w["test"].shape[0]
    row_bytes = 4 * (I + 1)
    fwd = 2 * (I * H + H * O) + (H + O)
    sgd = fwd + 4 * H * O + 2 * I * H + 5 * H + 5 * O
    sig = H + O
    trace = (w["P"] + 4) * 4
    rw = dict(flop=(N + M) * fwd, bytes=(N + M) * row_bytes + trace, sfu=(N + M) * sig * 2)
    lg = dict(flop=2 * N * sgd + (N + M) * fwd, bytes=(3 * N + M) * row_bytes + trace, sfu=(3 * N + M) * sig * 2)
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
            "vs_baseline": None, "dtype": "f64", "data": "synthetic" if w["name"] != "sunspot" else "Sunspot (reference dataset), random-init weights",
            "config": {"workload": w["desc"], "sampler": "Langevin PT, l_prob %.2f, lr %g, swap_interval %d, maxtemp %s" % (w["l_prob"], w["learn_rate"], w["swap_interval"], w["maxtemp"])},
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


def run_b200_arm(args, w, rank, world, local_rank):
    import torch
    from ptnn_b200 import capi
    from ptnn_b200.sampler import Sampler
    if capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libptfnn has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    K, W, si = args.steps, args.warmup, w["swap_interval"]
    S = si * (K + W) + 2
    stream = torch.cuda.current_stream(dev)
    kw = dict(use_langevin_gradients=True, l_prob=w["l_prob"], learn_rate=w["learn_rate"], seed=args.seed,
              common_random_numbers=True)
    rs = np.random.RandomState(1000 + rank)

    def make(memo, samples):
        if world > 1:
            from ptnn_b200.distributed import make_gpu_ladder
            ladder, smp = make_gpu_ladder(w["task"], w["topology"], w["temperatures"], samples, si, device=local_rank,
                                          memoize_gradient=memo, **kw)
        else:
            smp = Sampler(w["task"], w["topology"], w["temperatures"], samples, si, device=local_rank,
                          memoize_gradient=memo, stream=stream, **kw)
            ladder = None
        smp.set_data(w["train"], w["test"])
        smp.init_chains(rs.randn(smp.R, smp.P))
        return smp, ladder

    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_steps(smp, ladder, n_warm, n_timed):
        adv = (lambda: ladder.run(si)) if ladder is not None else (lambda: smp.run(si))
        for _ in range(n_warm):
            adv()
        barrier()
        ev, t_wall0 = [], time.perf_counter()
        for _ in range(n_timed):
            flush.fill_(1)                                   # evict L2 between timed steps (untimed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            adv()
            b.record(stream)
            ev.append((a, b))
        barrier()
        wall = time.perf_counter() - t_wall0
        ms = [a.elapsed_time(b) for a, b in ev]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)       # max over ranks
        return float(tot.item()), ms, wall

    # ---- (1) device-resident throughput, every logical pass executed (memoize_gradient = 0)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    smp, ladder = make(0, S)
    total_ms, ms_list, _ = timed_steps(smp, ladder, W, K)
    clk = clocks.stop() if clocks else None
    units_per_step = w["R_global"] * si
    value = units_per_step * K / (total_ms * 1e-3)
    # exact Langevin / random-walk mix of the timed steps (common random numbers: one lx per step)
    lx = smp.generate_draws(W * si, K * si)[0][0]
    n_lg = int(np.sum(lx < w["l_prob"]))
    n_rw = K * si - n_lg
    smp.close()

    # ---- (2) the same with the product default (memoised langevin_gradient(w); identical results)
    smp, ladder = make(1, S)
    total_ms_memo, _, _ = timed_steps(smp, ladder, W, K)
    value_memo = units_per_step * K / (total_ms_memo * 1e-3)
    smp.close()

    # ---- (3) end to end through the public API with host buffers: every step uploads the dataset
    #          (pinned-size host arrays) and reads the step's trace rows back
    smp, ladder = make(0, S)
    adv = (lambda: ladder.run(si)) if ladder is not None else (lambda: smp.run(si))
    for _ in range(W):
        adv()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for k in range(K):
        smp.set_data(w["train"], w["test"])                 # H2D
        adv()
        first = smp.step - si + 1
        t = smp.traces(first=first, count=si, pos_w=True, debug=False)   # D2H (synchronises)
        d2h = sum(v.nbytes // 2 if v.dtype == np.float64 and kname == "pos_w" else v.nbytes for kname, v in t.items())
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    I = w["topology"][0]
    ip = (I + 3) & ~3
    h2d = sum((n * ip + ((n + 3) & ~3)) * 4 for n in (w["train"].shape[0], w["test"].shape[0]))
    e2e_value = units_per_step * K / e2e_s
    # ---- (4) result pipeline on the device traces: burn-in slice -> pooled statistics, HBM-bound.
    #          Event-timed inside the library on the handle's stream; L2 flushed before every call.
    pipe = None
    if rank == 0:
        n_rows = smp.step
        smp.trace_summary(1, n_rows)
        got = []
        for _ in range(3):
            flush.fill_(1)
            torch.cuda.synchronize(dev)
            got.append(smp.trace_summary(1, n_rows))
        ms = float(np.mean([g["kernel_ms"] for g in got]))
        pipe = {"kernel": "trace_summary_kernel<2>", "bound": "hbm", "rows_pooled": int(got[0]["n"]),
                "algorithmic_bytes_per_launch": int(got[0]["bytes_read"]), "launch_ms": ms,
                "achieved": got[0]["bytes_read"] / (ms * 1e-3) / 1e9, "unit": "GB/s"}
    smp.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel in the timed region: chain_kernel
    rw, lg = algorithmic_work(w)
    Rg = w["R_global"]
    flops = Rg * (n_lg * lg["flop"] + n_rw * rw["flop"])
    byts = Rg * (n_lg * lg["bytes"] + n_rw * rw["bytes"])
    sfu = Rg * (n_lg * lg["sfu"] + n_rw * rw["sfu"])
    sec = total_ms * 1e-3
    peaks = measure_peaks(local_rank) or {}
    mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(mp_path):
        hbm_peak, hbm_src = float(json.load(open(mp_path))["hbm_gbs"]), "measured"
    fp32_peak = peaks.get("fp32_tflops")
    ach_tflops = flops / sec / 1e12 / world                 # per GPU
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(w["name"])
    fp32_roof = {"kernel": "chain_kernel<%d,%d,%d>" % w["topology"], "bound": "fp32-issue (not HBM: datasets are SMEM/L2 resident, SURVEY 8d)",
                 "achieved": ach_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": (ach_tflops / fp32_peak) if fp32_peak else None,
                 "peak_source": "measured in this run (csrc/ptfnn_peaks.cu FMA microbenchmark)", "traffic": traffic,
                 "algorithmic_flop_per_launch": flops / K / world, "launch_ms": total_ms / K}
    # the bounding on-chip resources are FP32 issue and the MUFU (XU) pipe; the one with the larger
    # fraction is reported as THE roofline (ncu agrees: XU is the busiest pipe of the bench launch), the other in roofline_alt
    mufu_peak = peaks.get("mufu_gops")
    ach_mufu = sfu / sec / 1e9 / world
    mufu_roof = {"kernel": "chain_kernel<%d,%d,%d>" % w["topology"], "bound": "MUFU / XU pipe: 2 transcendental ops per sigmoid (not HBM: datasets are SMEM/L2 resident, SURVEY 8d)",
                 "achieved": ach_mufu / 1e3, "peak": (mufu_peak / 1e3) if mufu_peak else None, "unit": "Tops/s",
                 "frac": (ach_mufu / mufu_peak) if mufu_peak else None,
                 "peak_source": "measured in this run (csrc/ptfnn_peaks.cu MUFU microbenchmark: 16 lanes/clk/SM)", "traffic": traffic,
                 "algorithmic_sfu_ops_per_launch": sfu / K / world, "launch_ms": total_ms / K}
    use_mufu = bool(mufu_roof["frac"] and fp32_roof["frac"] and mufu_roof["frac"] > fp32_roof["frac"])
    roofline = mufu_roof if use_mufu else fp32_roof
    roofline_alt = {
        "hbm": {"bound": "hbm", "achieved": byts / sec / 1e9 / world, "peak": hbm_peak, "unit": "GB/s",
                "frac": byts / sec / 1e9 / world / hbm_peak, "peak_source": hbm_src + " (MEASURED_PEAKS.json)" if hbm_src == "measured" else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": byts / K / world},
        ("fp32_issue" if use_mufu else "mufu"): (fp32_roof if use_mufu else mufu_roof),
        "smem_broadcast": {"achieved_gbs": byts / sec / 1e9 / world, "peak_gbs": peaks.get("smem_gbs")},
    }
    # The Langevin steps are a serial recurrence over the training rows: their bound is the
    # dependent-issue latency of one row, not a throughput peak.  floor = sum of the measured latencies of the
    # instructions on the chain (tools/latency_probe.cu); the upper bound charges the WHOLE
    # timed region (random-walk steps included) to the serial rows.
    chain_floor = {(4, 64, 1): 176, (4, 5, 1): 138, (4, 10, 1): 150, (16, 256, 10): 346}.get(tuple(w["topology"]))
    serial_rows = n_lg * 2 * w["train"].shape[0]
    if chain_floor and serial_rows and clk and clk.get("sm_mhz"):
        cyc = sec / serial_rows * clk["sm_mhz"] * 1e6
        roofline_alt["serial_chain"] = {"bound": "dependent-issue latency of the SGD recurrence (one chain per temperature)",
                                        "serial_rows_per_temperature": serial_rows, "cycles_per_row_upper_bound": cyc,
                                        "floor_cycles_per_row": chain_floor, "frac": chain_floor / cyc}

    # ---- CPU baseline on this box's host cores (bounded sample, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        cores = cpu_baseline.host_cores()
        n_procs, pattern = cpu_sample_plan(w, cores)
        wall, units = cpu_sample(w, n_procs, pattern)
        cpu = {"value": units / wall, "unit": "replica-steps/s", "cores": cores, "kind": "port",
               "sample": "%d replica processes x %d steps (LG/RW alternating = l_prob 0.5) of the same workload, %.1f s wall" % (n_procs, len(pattern), wall)}

    line = {"metric": "replica_mcmc_steps_per_sec", "value": value, "unit": "replica-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic" if w["name"] != "sunspot" else "Sunspot (reference dataset), random-init weights",
            "config": {"workload": w["desc"], "sampler": "Langevin PT, l_prob %.2f, lr %g, swap_interval %d, maxtemp %s" % (w["l_prob"], w["learn_rate"], si, w["maxtemp"]),
                       "replicas_total": Rg, "replica_steps_per_bench_step": units_per_step,
                       "langevin_steps_in_timed_region": n_lg, "random_walk_steps_in_timed_region": n_rw,
                       "memoize_gradient": 0, "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": "ladder partitioned over %d GPU(s), swap rounds in-kernel over NVLink peer memory" % world if world > 1 else "1 GPU, in-kernel swap round"},
            "value_memoized": value_memo, "ms_per_step_memoized": total_ms_memo / K,
            "e2e": {"value": e2e_value, "unit": "replica-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": K, "clocks": clk, "roofline": roofline, "roofline_alt": roofline_alt,
            "cpu_baseline": cpu, "peaks_measured": peaks, "step_ms": ms_list}
    if pipe:
        pipe.update(peak=hbm_peak, frac=pipe["achieved"] / hbm_peak, peak_source=roofline_alt["hbm"]["peak_source"])
        line["result_pipeline"] = pipe
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    w = load_workload(args.workload, n_dev, args.replicas_per_gpu)
    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, w)
        return
    run_b200_arm(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
