"""Read sharding behind the C ABI (BASELINE config 3): one handle that cuts every batch into one shard per
GPU (drprg_cuda_index_load_multi) and the one-process-per-GPU form (drprg_cuda_shard_root / attach / done).  Every
non-root shard adds its coverage straight into the root's accumulator (red.global.add over peer-mapped memory), so the
result must be bit-identical to the oracle's for any number of shards.  A device may be listed more than once, which is
how a one-GPU box exercises the whole sharded path (host threads, shard uploads, remote adds, hit listing)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as O
from drprg_b200 import lib, sim
from helpers import TOY_PRG, TOY_REFS, long_reads_sample, panel_sample, small_panel
from test_gpu_parity import assert_genotype_equal, assert_map_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    return int(lib.lib().drprg_cuda_device_count())


def run_sharded(gx, ox, data, off, refs, illumina=True, genome_size=4411532, stride_words=0, c=10):
    oo = O.make_opts(illumina=illumina, genome_size=genome_size, min_cluster_size=c)
    go = lib.make_opts(illumina=illumina, genome_size=genome_size, min_cluster_size=c)
    mr = O.MapRun(ox, data, off, oo)
    words, woff, lens = lib.pack_reads(data, off, stride_words)
    gx.sample_begin(go, int(off[1] - off[0]))
    nh, nk = gx.map_batch(gx.upload(words, woff, lens, total_bases=int(off[-1]), stride_words=stride_words))
    gh = gx.last_hits(nh)
    assert int(gh["kept"].sum()) == nk
    assert_map_equal(gx, mr, gh)
    assert_genotype_equal(gx, ox, mr, oo, refs)
    return nh


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0, 0, 0]])
def test_sharded_handle_shards_sharing_one_gpu(devices):
    p, prg, refs = small_panel()
    gx = lib.Index(prg, 11, 15, devices=devices)
    ox = O.Index(prg, 11, 15)
    assert gx.n_gpus == len(devices)
    d, o, g, pl = panel_sample(p, 50_000, seed=51)
    assert run_sharded(gx, ox, d, o, refs, genome_size=len(g), stride_words=10) > 10_000
    # a second sample on the same handle, ragged long reads (no -I), fewer reads than would fill every shard evenly
    d, o, g, pl = long_reads_sample(p, 203, seed=52)
    run_sharded(gx, ox, d, o, refs, illumina=False, genome_size=len(g))
    # more shards than reads
    d, o = sim.toy_dataset(TOY_PRG, TOY_REFS, depth=1, decoys=0, seed=3)
    gt, ot = lib.Index(TOY_PRG, 11, 15, devices=devices), O.Index(TOY_PRG, 11, 15)
    run_sharded(gt, ot, d[:int(o[3])], o[:4], TOY_REFS, genome_size=2000, stride_words=10)


def test_sharded_handle_drop_in_call(tmp_path):
    """the reference's one blocking call on a multi-shard handle: files in, pandora_genotyped.vcf out"""
    p, prg, refs = small_panel()
    d, o, g, pl = panel_sample(p, 30_000, seed=53)
    fq = tmp_path / "reads.fq"
    sim.write_fastq(str(fq), d, o)
    gx = lib.Index(prg, 11, 15, devices=[0, 0, 0])
    st = gx.map_genotype(fq, refs, tmp_path, lib.make_opts(illumina=True, genome_size=len(g), threads=4))
    ox = O.Index(prg, 11, 15)
    oo = O.make_opts(illumina=True, genome_size=len(g))
    og = O.Genotype(ox, O.MapRun(ox, d, o, oo), oo, refs)
    strip = lambda t: [l for l in t.splitlines() if not l.startswith("##fileDate")]
    assert strip((tmp_path / "pandora_genotyped.vcf").read_text()) == strip(og.vcf())
    assert st["n_reads"] == len(o) - 1


@pytest.mark.skipif(device_count() < 2, reason="needs two GPUs")
def test_sharded_handle_on_distinct_gpus():
    p, prg, refs = small_panel()
    n = min(device_count(), 8)
    gx = lib.Index(prg, 11, 15, n_gpus=n)
    ox = O.Index(prg, 11, 15)
    d, o, g, pl = panel_sample(p, 80_000, seed=54)
    run_sharded(gx, ox, d, o, refs, genome_size=len(g), stride_words=10)


@pytest.mark.skipif(device_count() < 2, reason="needs two GPUs")
def test_one_process_per_gpu_fused_reduce():
    """torchrun, 2 ranks: accumulators and VCF of the fused (peer-memory) reduce == NCCL allreduce == one process"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "tools", "sharded_parity.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded parity ok" in r.stdout
