"""Summarise an .ncu-rep (raw page -> key metrics; source page -> loop instruction mix and stalls).
Usage: python tools/ncu_summary.py gpurun_out/force_x.ncu-rep [kernel-index]"""
import collections
import csv
import io
import re
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = r[0], r[1], r[2:]
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    keys += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for row in rows:
        print("-" * 100)
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                name = k.replace("smsp__average_warps_issue_stalled_", "stall:").replace("_per_issue_active.ratio", "")
                print("%-75s %s %s" % (name, row[i], units[i]))


def source(rep, which=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        return
    s = starts[min(which, len(starts) - 1)]
    e = starts[which + 1] if which + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[s + 2:e] if len(r) >= len(hdr)]
    mx = max(int(r[idx["Instructions Executed"]]) for r in data)
    loop = [r for r in data if int(r[idx["Instructions Executed"]]) >= 0.9 * mx]
    ops = collections.Counter()
    for r in loop:
        m = re.match(r"(@!?U?P\d\s+)?(\S+)", r[1].strip())
        ops[m.group(2).split(".")[0]] += 1
    tot = sum(int(r[idx["# Samples"]]) for r in loop)
    print("=" * 100)
    print(rows[s][1][:110])
    print("hot loop: %d instructions, executed %d times each; %d of %d samples" % (
        len(loop), mx, tot, sum(int(r[idx["# Samples"]]) for r in data)))
    print("instruction mix:", dict(ops.most_common()))
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[idx[c]]) for r in loop) / max(tot, 1) for c in cols}
    print("stall shares in loop:", {k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005})


if __name__ == "__main__":
    raw(sys.argv[1])
    source(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
