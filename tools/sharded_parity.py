"""Read-sharded run under torchrun (NCCL) against the same reads mapped by one process: accumulators and VCF must be
identical for any world size.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_parity.py [n_reads]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
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
ix = lib.Index(wl.prg_path, wl.w, wl.k, device=local)
opts = lib.make_opts(illumina=True)
sub_off = o[lo:hi + 1] - o[lo]
words, _, lens = lib.pack_reads(d[int(o[lo]):int(o[hi])], sub_off, workload.STRIDE_WORDS)
ix.sample_begin(opts, workload.READ_LEN)
ix.map_batch(ix.upload(words, None, lens, total_bases=int(sub_off[-1]), stride_words=workload.STRIDE_WORDS, read_id_base=lo))
sharded.allreduce_accum(ix)
torch.cuda.synchronize()
acc = ix.accum_download()
ok = None
if rank == 0:
    ix.genotype(wl.refs_path)
    vcf_sharded = [l for l in ix.vcf().splitlines() if not l.startswith("##fileDate")]
    w2, _, l2 = lib.pack_reads(d, o, workload.STRIDE_WORDS)
    ix.sample_begin(opts, workload.READ_LEN)
    ix.map_batch(ix.upload(w2, None, l2, total_bases=int(o[-1]), stride_words=workload.STRIDE_WORDS))
    whole = ix.accum_download()
    ix.genotype(wl.refs_path)
    vcf_whole = [l for l in ix.vcf().splitlines() if not l.startswith("##fileDate")]
    # the four scalar words are lo24/hi partial sums: compare them decoded, everything else raw
    acc_eq = bool((acc[:-4] == whole[:-4]).all()) and sharded.decode_scalars(acc[-4:]) == sharded.decode_scalars(whole[-4:])
    ok = acc_eq and vcf_sharded == vcf_whole
    print(json.dumps({"world": world, "reads": n, "accumulators_equal": acc_eq, "vcf_equal": vcf_sharded == vcf_whole,
                      "vcf_lines": len(vcf_whole)}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok in (None, True) else 1)
