"""Instruction mix and hottest SASS lines of one kernel from `ncu -i rep --page source --csv` output.
usage: ncu_source_mix.py file.csv [top_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
si, ei, ti = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
agg, tot, lines = collections.Counter(), 0, []
for r in rows[hi + 1:]:
    if len(r) <= ei or r[0] in ("Kernel Name", "Address"):
        if r and r[0] == "Kernel Name":
            break
        continue
    try:
        e = int(r[ei])
    except ValueError:
        continue
    tot += e
    t = r[si].split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") else t[0]
    agg[op.split(".")[0]] += e
    lines.append((e, int(r[ti] or 0), r[si].strip()))
print("warp instructions executed:", tot)
for k, v in agg.most_common(top):
    print(f"{k:12s} {v:12d} {100 * v / tot:5.1f}%")
print("--- hottest lines (executed, samples)")
for e, s, src in sorted(lines, reverse=True)[:top]:
    print(f"{e:10d} {s:6d}  {src}")
