"""BASELINE config 5: a batch of simulated Illumina samples, SAMPLE-sharded over the GPUs of the box (replicas only: no
collective on the data path).  Rank r of N keeps the index resident on its GPU and runs samples r, r + N, ... through the
drop-in batch call drprg_cuda_map_genotype_batch (FASTQ files in, pandora_genotyped.vcf files out).  Reports samples/hour
end to end and checks every VCF against the VCF the same reads give on rank 0 alone.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/batch_bench.py [n_samples] [gz]"""
import ctypes as C
import hashlib
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from drprg_b200 import lib, sharded, sim, workload

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 96
gz = len(sys.argv) > 2 and sys.argv[2] == "gz"
cores = os.cpu_count() or 1
os.environ.setdefault("DRPRG_THREADS", str(max(2, cores // world)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
wl = workload.Config3()
DISTINCT = 4  # distinct read sets (different sub-shards of the workload = different read seeds); samples cycle through them
root = "/dev/shm" if os.path.isdir("/dev/shm") else None
tmp = os.path.join(root or tempfile.gettempdir(), "drprg_batch_bench")
if rank == 0:
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    for s in range(DISTINCT):
        codes = torch.cat([wl.subshard_codes(8 * s + i) for i in range(8)])  # 1 M reads
        data = sim.BASES[codes.cpu().numpy()].reshape(-1)
        off = np.arange(codes.shape[0] + 1, dtype=np.uint64) * np.uint64(workload.READ_LEN)
        sim.write_fastq_fast(os.path.join(tmp, f"sample{s}.fq" + (".gz" if gz else "")), data, off, gz=gz, seed=s)
if world > 1:
    dist.barrier()
ix = lib.Index(wl.prg_path, wl.w, wl.k, device=local)
opts = lib.make_opts(illumina=True, genome_size=workload.GENOME_SIZE, threads=max(2, cores // world))
mine = list(range(n_samples))[rank::world]
reads = [os.path.join(tmp, f"sample{s % DISTINCT}.fq" + (".gz" if gz else "")).encode() for s in mine]
outs = []
for s in mine:
    od = os.path.join(tmp, f"out{s}")
    os.makedirs(od, exist_ok=True)
    outs.append(od.encode())
# warm-up: one sample (buffers, site tables)
ix.map_genotype(reads[0].decode(), wl.refs_path, outs[0].decode(), opts)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
arr_r, arr_o = (C.c_char_p * len(mine))(*reads), (C.c_char_p * len(mine))(*outs)
rc = lib.lib().drprg_cuda_map_genotype_batch(ix.h, C.c_size_t(len(mine)), arr_r, wl.refs_path.encode(), arr_o, C.byref(opts), None)
assert rc == 0, lib.lib().drprg_cuda_last_error()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0


def sha(path):
    return hashlib.sha1(b"\n".join(l for l in open(path, "rb").read().splitlines() if not l.startswith(b"##fileDate"))).hexdigest()


ok = True
if rank == 0:
    want = {}
    for s in range(min(DISTINCT, n_samples)):
        od = os.path.join(tmp, f"ref{s}")
        os.makedirs(od, exist_ok=True)
        ix.map_genotype(os.path.join(tmp, f"sample{s}.fq" + (".gz" if gz else "")), wl.refs_path, od, opts)
        want[s] = sha(os.path.join(od, "pandora_genotyped.vcf"))
    for s in range(n_samples):
        ok = ok and sha(os.path.join(tmp, f"out{s}", "pandora_genotyped.vcf")) == want[s % DISTINCT]
    print(json.dumps({"config": "config5: batch of simulated Illumina samples (1 M x 150 bp reads each), sample-sharded, replicas only",
                      "n_gpus": world, "samples": n_samples, "input": "gzip FASTQ" if gz else "plain FASTQ", "seconds": dt,
                      "samples_per_hour": 3600.0 * n_samples / dt, "ms_per_sample_per_gpu": dt / max(1, len(mine)) * 1e3,
                      "every_vcf_equals_single_gpu_run": bool(ok), "host_cores": cores}))
    shutil.rmtree(tmp, ignore_errors=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
