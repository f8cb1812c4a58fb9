"""BASELINE config 3 from a FILE through the drop-in call: the first N reads of the 30 M-read workload are written as one
plain FASTQ (315 bytes per read: 10 M reads = 3.15 GB) and mapped + genotyped by drprg_cuda_map_genotype, which reads it
wave by wave (1 GiB of text per wave).  The VCF must equal the one the staged calls give for the same reads resident in
HBM.
   python tools/big_file_probe.py [n_reads=10000000] [dir=/dev/shm] [gz]      (gz: the file is compressed with `gzip -1` first)"""
import hashlib, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from drprg_b200 import lib, sim, workload

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
root = sys.argv[2] if len(sys.argv) > 2 else ("/dev/shm" if os.path.isdir("/dev/shm") else None)
wl = workload.Config3(total_reads=n_total)
tmp = tempfile.mkdtemp(prefix="drprg_big_", dir=root)
fq = os.path.join(tmp, "reads.fq")
L = workload.READ_LEN
t0 = time.perf_counter()
wl.write_fastq(fq)
if len(sys.argv) > 3 and sys.argv[3] == "gz":
    subprocess.run(f"gzip -1 -c {fq} > {fq}.gz && rm {fq}", shell=True, check=True)
    fq = fq + ".gz"
t_write = time.perf_counter() - t0
ix = lib.Index(wl.prg_path, wl.w, wl.k, device=0)
fo = lib.make_opts(illumina=True, genome_size=workload.GENOME_SIZE, threads=os.cpu_count() or 1)
ms = []
for _ in range(3):
    t0 = time.perf_counter()
    st = ix.map_genotype(fq, wl.refs_path, tmp, fo)
    ms.append((time.perf_counter() - t0) * 1e3)
strip = lambda b: b"\n".join(l for l in b.splitlines() if not l.startswith(b"##fileDate"))
sha_file = hashlib.sha1(strip(open(os.path.join(tmp, "pandora_genotyped.vcf"), "rb").read())).hexdigest()
log = open(os.path.join(tmp, "pandora.log")).read()
# the same reads resident in HBM through the staged calls
opts = lib.make_opts(illumina=True, genome_size=workload.GENOME_SIZE)
ix.sample_begin(opts, L)
group = 16
for g0 in range(0, wl.n_subshards, group):
    ids = range(g0, min(g0 + group, wl.n_subshards))
    w = torch.cat([wl.pack_codes(wl.subshard_codes(i)) for i in ids])
    l = torch.full((w.shape[0],), L, dtype=torch.int32, device="cuda")
    ix.map_batch(ix.wrap_device(w.data_ptr(), l.data_ptr(), w.shape[0], workload.STRIDE_WORDS, w.shape[0] * L,
                                read_id_base=g0 * wl.SUBSHARD, keep=(w, l)))
ix.genotype(wl.refs_path)
sha_res = hashlib.sha1(strip(bytes(ix.vcf_view()))).hexdigest()
print(json.dumps({"workload": f"config 3, first {n_total} reads, {'gzip -1' if fq.endswith('.gz') else 'plain'} FASTQ file ({os.path.getsize(fq) / 1e9:.2f} GB, page cache warm) -> pandora_genotyped.vcf",
                  "ms_per_call": [round(x, 1) for x in ms], "reads_per_s": n_total / (min(ms) * 1e-3), "n_reads": st["n_reads"],
                  "ingest_ms": round(st["ms_ingest"], 1), "map_ms": round(st["ms_map"], 1), "genotype_ms": round(st["ms_genotype"], 2),
                  "log_ingest": [x for x in log.split() if "wave" in x or "framed" in x][:3], "host_cores": os.cpu_count(),
                  "vcf_equals_resident_run": sha_file == sha_res, "write_s": round(t_write, 1)}))
import shutil
shutil.rmtree(tmp, ignore_errors=True)
