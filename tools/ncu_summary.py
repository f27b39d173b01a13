"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md: per-launch duration,
occupancy limits, pipe utilisation, stall reasons, DRAM traffic.  usage: ncu_summary.py rep out.md [title]"""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__icc_request_hit_rate.pct",
]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
with open(out, "w") as f:
    f.write("# %s\n\nsource: `%s` (ncu --set full --clock-control none --import-source on); %d launch(es)\n\n" % (title, rep, len(data)))
    f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
    for k in KEYS:
        if k in col:
            f.write("| %s | %s | %s |\n" % (k, units[col[k]], " | ".join(r[col[k]][:60] for r in data)))
    f.write("\n## warps stalled per issue, by reason\n\n| reason | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n|---|" + "---|" * len(data) + "\n")
    for h in sorted(STALL, key=lambda h: -max(float(r[col[h]] or 0) for r in data)):
        f.write("| %s | %s |\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                                   " | ".join(r[col[h]] for r in data)))
print("wrote", out)
