#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep and the ncu launch list into the small text summaries committed under profiles/.
Usage: python profiles/summarise.py <round-tag>   (reads gpurun_out/, writes profiles/<tag>_*.{md,csv})"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def summarise_rep(rep, tag, name):
    hdr, units, rows = raw_rows(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {name} — ncu --set full --clock-control none ({os.path.basename(rep)})", ""]
    for r in rows:
        lines.append(f"## {r[ix['Kernel Name']]}  (launch id {r[ix['ID']]})")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in ix:
                lines.append(f"| {k} | {r[ix[k]]} | {units[ix[k]]} |")
        lines.append("")
    open(os.path.join(OUT, f"{tag}_{name}.md"), "w").write("\n".join(lines))


def summarise_launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.defaultdict(list)
    order = []
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = r[ix["Kernel Name"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v = v / 1000.0 if unit in ("ns", "nsecond") else v
        if unit in ("ms", "msecond"):
            v *= 1000.0
        if k not in per:
            order.append(k)
        per[k].append(v)
    total = sum(sum(v) for v in per.values())
    lines = ["# launch list: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)",
             "", "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k in sorted(per, key=lambda k: -sum(per[k])):
        s = sum(per[k])
        lines.append(f"| `{k[:110]}` | {len(per[k])} | {s:.1f} | {s / len(per[k]):.1f} | {100 * s / total:.1f}% |")
    open(os.path.join(OUT, f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    for f in sorted(os.listdir(GP)):
        if f.endswith(".ncu-rep") and f.startswith(f"prof_{tag}_"):
            summarise_rep(os.path.join(GP, f), tag, f[len(f"prof_{tag}_"):-len(".ncu-rep")])
    ll = os.path.join(GP, f"launches_{tag}.csv")
    if os.path.exists(ll):
        summarise_launches(ll, tag)
