"""Summarise an ncu --metrics gpu__time_duration.sum --csv launch list: per kernel name, launches and mean/total microseconds
(optionally only the last N launches, i.e. the last steps of a bench run).  usage: launch_list.py file.csv [last_n]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H = rows[hdr]
rows = rows[hdr + 1:]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
seq = [(r[ki].split("(")[0].replace("void ", "").replace("drprg::", "")[:48], float(r[vi].replace(",", "")) / 1000.0) for r in rows]
if len(sys.argv) > 2:
    seq = seq[-int(sys.argv[2]):]
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':50s} {'n':>4s} {'mean us':>9s} {'total us':>9s} {'share':>6s}")
for k, (n, t) in agg.items():
    print(f"{k:50s} {n:4d} {t / n:9.1f} {t:9.1f} {100 * t / tot:5.1f}%")
