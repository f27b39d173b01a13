"""Print the headline and every sub-record of a bench.py JSON line in one screen.  usage: bench_digest.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def show(name, r):
    if not isinstance(r, dict) or "value" not in r:
        print(name, r)
        return
    tr = r.get("timed_region", {})
    acc = tr.get("acceptance_rate")
    cpu = (r.get("cpu_baseline") or {}).get("value")
    print("%-22s value %9.0f  memo %9.0f  e2e %9.0f (%.3f)  ms/step %7.2f  acc %s  LG/RW %s/%s  roof %.3f [%s]%s" % (
        name, r["value"], r["value_memoized"], r["e2e"]["value"], r["e2e"]["value"] / r["value"], r["ms_per_step"],
        "%.4f" % acc if acc is not None else "-", tr.get("langevin_steps"), tr.get("random_walk_steps"), r["roofline"]["frac"] or -1,
        r["roofline"]["bound"][:10], "  cpu %.1f (x%.0f, e2e x%.0f)" % (cpu, r["value"] / cpu, r["e2e"]["value"] / cpu) if cpu else ""))
    for k, v in r.get("roofline_alt", {}).items():
        if isinstance(v, dict) and v.get("frac") is not None:
            print("      alt %-12s frac %.3f" % (k, v["frac"]))
    if "burn_in" in r:
        print("      burn-in:", r["burn_in"])


show("head (n_gpus %d)" % d["n_gpus"], d)
for k, v in d.get("sub", {}).items():
    show(k, v)
if "result_pipeline" in d:
    print("result_pipeline frac %.3f (%.0f GB/s)" % (d["result_pipeline"]["frac"], d["result_pipeline"]["achieved"]))
if "multi_gpu_parity" in d:
    print("multi_gpu_parity", d["multi_gpu_parity"], d.get("multi_gpu_parity_detail"))
print("clocks", d.get("clocks"))
