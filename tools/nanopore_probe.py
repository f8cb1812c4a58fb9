"""BASELINE config 4 shape: the config-2 panel, simulated ~10 kb reads at 5 % error (no -I), timed through the C ABI;
a slice is checked against the oracle."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from drprg_b200 import lib, sim, workload
wl = workload.Config2()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15000
t = time.time(); d, o = sim.simulate_long_reads(wl.genome, n, mean_len=10000, seed=77); print("simulated", n, "reads", int(o[-1]), "bases in %.1fs" % (time.time() - t))
words, woff, lens = lib.pack_reads(d, o)
ix = lib.Index(wl.prg_path, 11, 15); opts = lib.make_opts(illumina=False)
b = ix.upload(words, woff, lens, total_bases=int(o[-1]))
for i in range(4):
    ix.sample_begin(opts, int(o[1] - o[0])); t0 = time.perf_counter(); nh, nk = ix.map_batch(b); t1 = time.perf_counter(); ix.genotype(wl.refs_path); t2 = time.perf_counter()
    print("map %.2f ms genotype %.2f ms hits %d kept %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, nh, nk), ix.last_timings())
print(json.dumps({"config": "config4 nanopore", "reads": n, "bases": int(o[-1]), "reads_per_s": n / (t2 - t0), "bases_per_s": int(o[-1]) / (t2 - t0), "records": len(ix.gt_records()["pos"])}))
import oracle_py as O
m = min(n, 1500)
ox = O.Index(wl.prg_path, 11, 15); oo = O.make_opts(illumina=False, threads=os.cpu_count())
t0 = time.perf_counter(); mr = O.MapRun(ox, d[: int(o[m])], o[: m + 1], oo); dt = time.perf_counter() - t0
print("oracle (%d threads): %d reads in %.2fs = %.0f reads/s" % (os.cpu_count(), m, dt, m / dt))
w2, o2, l2 = lib.pack_reads(d[: int(o[m])], o[: m + 1])
ix.sample_begin(opts, int(o[1] - o[0])); nh, nk = ix.map_batch(ix.upload(w2, o2, l2, total_bases=int(o[m])))
gh, oh = ix.last_hits(nh), mr.hits()
ok = all(len(gh[k]) == len(oh[k]) and (gh[k] == oh[k]).all() for k in ("read", "prg", "fwd", "start", "knode", "kept"))
f, r = mr.coverage(); cov = ix.coverage()
print("parity on %d reads: hits %s coverage %s" % (m, ok, bool((cov["fwd"] == f).all() and (cov["rev"] == r).all())))
