"""BASELINE config 5 shape on one GPU: a batch of samples (1 M x 150 bp reads each, different read seeds) through
drprg_cuda_map_genotype_batch, files in, VCFs out; samples/hour for plain and gzip FASTQ.
   python tools/batch_probe.py [n_samples] [reads_per_sample]"""
import ctypes as C, os, subprocess, sys, tempfile, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drprg_b200 import lib, workload

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
wl = workload.Config2()
tmp = tempfile.mkdtemp()
L = 150
paths = []
for s in range(ns):
    d, o = wl.reads(n, s)
    rec = np.empty((n, 8 + 1 + L + 3 + L + 1), np.uint8)
    ids = np.char.zfill(np.arange(n).astype(str), 7).astype("S7")
    rec[:, 0] = ord("@"); rec[:, 1:8] = np.frombuffer(ids.tobytes(), np.uint8).reshape(n, 7); rec[:, 8] = 10
    rec[:, 9:9 + L] = d.reshape(n, L); rec[:, 9 + L] = 10; rec[:, 10 + L] = ord("+"); rec[:, 11 + L] = 10
    rec[:, 12 + L:12 + 2 * L] = ord("I"); rec[:, 12 + 2 * L] = 10
    fq = os.path.join(tmp, f"s{s}.fq"); rec.tofile(fq); paths.append(fq)
procs = [subprocess.Popen(f"gzip -1 -c {p} > {p}.gz", shell=True) for p in paths]
[p.wait() for p in procs]
ix = lib.Index(wl.prg_path, 11, 15); opts = lib.make_opts(illumina=True, threads=os.cpu_count() or 1)
Lb = lib.lib()
out = {}
for kind, files in (("plain", paths), ("gzip", [p + ".gz" for p in paths])):
    outs = []
    for s in range(ns):
        od = os.path.join(tmp, f"{kind}{s}"); os.makedirs(od, exist_ok=True); outs.append(od.encode())
    arr_r = (C.c_char_p * ns)(*[f.encode() for f in files]); arr_o = (C.c_char_p * ns)(*outs)
    stats = (lib.MapStats * ns)()
    best = None
    for rep in range(2):
        t0 = time.perf_counter()
        rc = Lb.drprg_cuda_map_genotype_batch(ix.h, C.c_size_t(ns), arr_r, wl.refs_path.encode(), arr_o, C.byref(opts), stats)
        dt = time.perf_counter() - t0
        assert rc == 0, Lb.drprg_cuda_last_error()
        best = dt if best is None else min(best, dt)
    out[kind] = dict(samples=ns, reads_per_sample=n, seconds=round(best, 4), samples_per_hour=round(ns / best * 3600), reads_per_s=round(ns * n / best))
print(json.dumps(out))
