"""profiles/r2_screen.json + .md from an `ncu --set full` capture of screen_kernel / resolve_kernel on a 1 M-read batch:
   ncu -i rep --page raw --csv > raw.csv ; python tools/make_screen_profile.py raw.csv n_reads out_prefix [how]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_reads = int(sys.argv[2])
out = sys.argv[3]
H, U = rows[0], rows[1]


def col(name):
    return H.index(name)


def num(r, name):
    v = float(r[col(name)].replace(",", ""))
    u = U[col(name)]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u, 1)


per = {}
for r in rows[2:]:
    name = r[col("Kernel Name")]
    key = "screen" if "screen_kernel" in name else "resolve" if "resolve_kernel" in name else None
    if not key or key in per:
        continue
    per[key] = {
        "kernel": name.split("(")[0],
        "time_us": num(r, "gpu__time_duration.sum") * 1e6,
        "warp_instr": num(r, "smsp__inst_executed.sum"),
        "dram_read_bytes": num(r, "dram__bytes_read.sum"),
        "dram_write_bytes": num(r, "dram__bytes_write.sum"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in H else None,
        "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") if "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active" in H else None,
        "fmaheavy_pipe_pct": num(r, "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active") if "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active" in H else None,
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": num(r, "launch__registers_per_thread"),
        "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct") if "lts__t_sector_hit_rate.pct" in H else None,
    }
scale = 1e6 / n_reads
d = {
    "reads_in_capture": n_reads,
    "warp_instr_per_read_screen": per["screen"]["warp_instr"] / n_reads,
    "warp_instr_per_read_resolve": per["resolve"]["warp_instr"] / n_reads,
    "dram_bytes_per_million_reads": (per["screen"]["dram_read_bytes"] + per["screen"]["dram_write_bytes"] + per["resolve"]["dram_read_bytes"] + per["resolve"]["dram_write_bytes"]) * scale,
    "kernels": per,
    "how": sys.argv[4] if len(sys.argv) > 4 else "ncu --set full --clock-control none on one launch of each kernel (cold caches, serialised); see profiles/README.md",
}
json.dump(d, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write(f"# sketch+lookup kernels, ncu --set full, one launch each on {n_reads} x 150 bp reads\n\n")
    f.write("| kernel | time us | warp-instr / read | DRAM read MB | DRAM write MB | issue active % | ALU pipe % | FMA-heavy pipe % | warps active % | regs |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for k, p in per.items():
        f.write(f"| {p['kernel']} | {p['time_us']:.1f} | {p['warp_instr'] / n_reads:.1f} | {p['dram_read_bytes'] / 1e6:.1f} | {p['dram_write_bytes'] / 1e6:.1f} | "
                f"{p['issue_active_pct']} | {p['alu_pipe_pct']} | {p['fmaheavy_pipe_pct']} | {p['warps_active_pct']:.1f} | {p['registers']:.0f} |\n")
    f.write(f"\nDRAM traffic of the pair: {d['dram_bytes_per_million_reads'] / 1e6:.1f} MB per million reads "
            f"(algorithmic: 44 MB of packed reads + 16 B per hit).\n")
print(json.dumps(d)[:400])
