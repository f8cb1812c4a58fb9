"""Read-sharded run with one process per GPU (torchrun) against the same reads mapped by one process: the fused reduce
(peer-memory adds into the root's accumulator), the NCCL allreduce and the single-process run must give identical
accumulators and VCF for any world size.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_parity.py [n_reads]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from drprg_b200 import lib, sharded, workload

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("DRPRG_THREADS", str(sharded.host_threads_for_rank(rank, world)))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
wl = workload.Config2()
d, o = wl.reads(n, 0)                      # every rank generates the same reads and takes its slice
lo, hi = sharded.shard_bounds(n, world, rank)
opts = lib.make_opts(illumina=True)
sub_off = o[lo:hi + 1] - o[lo]
words, _, lens = lib.pack_reads(d[int(o[lo]):int(o[hi])], sub_off, workload.STRIDE_WORDS)
strip = lambda t: [l for l in t.splitlines() if not l.startswith("##fileDate")]


def shard_batch(ix):
    return ix.upload(words, None, lens, total_bases=int(sub_off[-1]), stride_words=workload.STRIDE_WORDS, read_id_base=lo)


# ---- (1) NCCL allreduce of the accumulator
ix = lib.Index(wl.prg_path, wl.w, wl.k, device=local)
ix.sample_begin(opts, workload.READ_LEN)
ix.map_batch(shard_batch(ix))
sharded.allreduce_accum(ix)
torch.cuda.synchronize()
acc_nccl = ix.accum_download()
vcf_nccl = None
if rank == 0:
    ix.genotype(wl.refs_path)
    vcf_nccl = strip(ix.vcf())
dist.barrier()
# ---- (2) fused reduce: two samples in a row on fresh handles (the second one checks the epoch / arrival bookkeeping)
fx = lib.Index(wl.prg_path, wl.w, wl.k, device=local)
sharded.setup_fused_reduce(fx, rank, world)
acc_fused, vcf_fused = [], []
for sample in range(2):
    fx.sample_begin(opts, workload.READ_LEN)
    fx.map_batch(shard_batch(fx))
    if rank == 0:
        fx.genotype(wl.refs_path)          # waits on the device for the other ranks' arrivals
        vcf_fused.append(strip(fx.vcf()))
        acc_fused.append(fx.accum_download())
    else:
        fx.shard_done()
ok = None
if rank == 0:
    w2, _, l2 = lib.pack_reads(d, o, workload.STRIDE_WORDS)
    sx = lib.Index(wl.prg_path, wl.w, wl.k, device=local)
    sx.sample_begin(opts, workload.READ_LEN)
    sx.map_batch(sx.upload(w2, None, l2, total_bases=int(o[-1]), stride_words=workload.STRIDE_WORDS))
    sx.genotype(wl.refs_path)
    whole = sx.accum_download()
    vcf_whole = strip(sx.vcf())

    def same(acc):  # the four scalar words are lo24/hi partial sums: compare them decoded, everything else raw
        return bool((acc[:-4] == whole[:-4]).all()) and sharded.decode_scalars(acc[-4:]) == sharded.decode_scalars(whole[-4:])

    res = {"world": world, "reads": n, "nccl_accumulators_equal": same(acc_nccl), "nccl_vcf_equal": vcf_nccl == vcf_whole,
           "fused_accumulators_equal": [same(a) for a in acc_fused], "fused_vcf_equal": [v == vcf_whole for v in vcf_fused],
           "vcf_lines": len(vcf_whole)}
    ok = res["nccl_accumulators_equal"] and res["nccl_vcf_equal"] and all(res["fused_accumulators_equal"]) and all(res["fused_vcf_equal"])
    print(json.dumps(res))
    if ok:
        print("sharded parity ok")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok in (None, True) else 1)
