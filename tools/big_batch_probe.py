"""One large batch (default 6 M reads, BASELINE config 3's per-GPU share is 3.75 M) vs the same reads in 1 M-read shards:
accumulators must be identical (exercises hit-buffer regrowth, chunked upload and 32-bit bookkeeping at scale)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from drprg_b200 import lib, workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
wl = workload.Config2()
parts = [wl.reads(1_000_000, s) for s in range(n // 1_000_000)]
words = np.concatenate([lib.pack_reads(d, o, 10)[0] for d, o in parts]); lens = np.full(n, 150, np.uint32)
ix = lib.Index(wl.prg_path, 11, 15); opts = lib.make_opts(illumina=True)
ix.sample_begin(opts, 150); t0 = time.perf_counter(); nh, nk = ix.map_batch(ix.upload(words, None, lens, total_bases=150 * n, stride_words=10)); t1 = time.perf_counter()
whole = ix.accum_download().astype(np.int64); ix.genotype(wl.refs_path); t2 = time.perf_counter()
print("one batch: %d reads, %d hits, %d kept; map %.2f ms (%s) genotype %.2f ms -> %.1f M reads/s" % (n, nh, nk, (t1 - t0) * 1e3, ix.last_timings(), (t2 - t1) * 1e3, n / (t2 - t0) / 1e6))
v1 = ix.vcf_bytes()
ix.sample_begin(opts, 150)
for s in range(n // 1_000_000):
    w = words[s * 10_000_000:(s + 1) * 10_000_000]
    ix.map_batch(ix.upload(w, None, lens[:1_000_000], total_bases=150_000_000, stride_words=10, read_id_base=s * 1_000_000))
sh = ix.accum_download().astype(np.int64); ix.genotype(wl.refs_path)
print("sharded == whole:", bool((sh == whole).all()), "vcf equal:", ix.vcf_bytes().split(b"\n", 3)[3] == v1.split(b"\n", 3)[3], "E =", ix.params()["E"])
