"""BASELINE.json configs 3 and 4 at their full sizes, CUDA path (through the C ABI) against the CPU oracle.

config 3: the config-2 panel and genome, read-sharded: 8 shards x 1 M simulated 150 bp reads, each mapped as its own
          batch into its own accumulator (what one rank does), the accumulators summed (what the reduce does), then the
          genotype step — coverage, locus read counts and the VCF must equal the oracle run on all 8 M reads at once.
config 4: the same panel in nanopore mode (no -I): 15 000 simulated reads of ~10 kb at 5 % error (40/30/30
          sub/ins/del) — hits, kept flags, coverage, parameters, ML paths, genotype arrays and VCF against the oracle."""
import os

import numpy as np
import pytest

import oracle_py as O
from drprg_b200 import lib, sim, workload
from test_gpu_parity import assert_genotype_equal, assert_map_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2():
    wl = workload.Config2()
    return wl, lib.Index(wl.prg_path, wl.w, wl.k, device=0), O.Index(wl.prg_path, wl.w, wl.k)


def test_config4_nanopore_15000_reads_of_10kb(cfg2):
    wl, gx, ox = cfg2
    d, o = sim.simulate_long_reads(wl.genome, 15_000, mean_len=10_000, sigma=0.3, seed=workload.PANEL_SEED + 3, err=0.05)
    assert 140e6 < int(o[-1]) < 160e6  # ~34x of the 4.4 Mb genome
    threads = os.cpu_count() or 1
    oo = O.make_opts(threads=threads, illumina=False, genome_size=workload.GENOME_SIZE)
    go = lib.make_opts(illumina=False, genome_size=workload.GENOME_SIZE)
    mr = O.MapRun(ox, d, o, oo)
    words, woff, lens = lib.pack_reads(d, o)
    gx.sample_begin(go, int(o[1] - o[0]))
    nh, nk = gx.map_batch(gx.upload(words, woff, lens, total_bases=int(o[-1])))
    gh = gx.last_hits(nh)
    assert nh > 100_000 and nk > 50_000 and int(gh["kept"].sum()) == nk
    assert_map_equal(gx, mr, gh)
    og = assert_genotype_equal(gx, ox, mr, oo, wl.refs_path)
    assert len(og.records()["pos"]) > 4000


def test_config3_eight_shards_of_1m_reads_summed(cfg2):
    wl, gx, ox = cfg2
    n_shards, per = 8, 1_000_000
    go = lib.make_opts(illumina=True, genome_size=workload.GENOME_SIZE)
    tot = np.zeros(gx.n_accum, np.int64)
    datas = []
    for s in range(n_shards):
        d, o = wl.reads(per, s)  # seed + 2 + shard
        datas.append(d)
        words, _, lens = lib.pack_reads(d, o, workload.STRIDE_WORDS)
        gx.sample_begin(go, workload.READ_LEN)
        gx.map_batch(gx.upload(words, None, lens, total_bases=int(o[-1]), stride_words=workload.STRIDE_WORDS, read_id_base=s * per))
        tot += gx.accum_download().astype(np.int64)
    # scalars travel as lo24/hi pairs: renormalise the sums like the reduce's consumer does
    tb = int(tot[-4]) + (int(tot[-3]) << 24)
    nr = int(tot[-2]) + (int(tot[-1]) << 24)
    assert tb == n_shards * per * workload.READ_LEN and nr == n_shards * per
    tot[-4:] = [tb & 0xFFFFFF, tb >> 24, nr & 0xFFFFFF, nr >> 24]
    d_all = np.concatenate(datas)
    o_all = np.arange(n_shards * per + 1, dtype=np.uint64) * np.uint64(workload.READ_LEN)
    del datas
    oo = O.make_opts(threads=os.cpu_count() or 1, illumina=True, genome_size=workload.GENOME_SIZE)
    mr = O.MapRun(ox, d_all, o_all, oo)
    f, r = mr.coverage()
    N = ox.total_knodes
    assert (tot[:2 * N:2] == f).all() and (tot[1:2 * N:2] == r).all()
    assert (tot[2 * N:2 * N + ox.n_loci] == mr.locus_reads()).all()
    gx.sample_begin(go, workload.READ_LEN)
    gx.accum_upload(tot.astype(np.int32))
    assert_genotype_equal(gx, ox, mr, oo, wl.refs_path)
