"""the mapping front half of `pandora discover` (/root/reference/src/predict.rs:247-256) served from the
map pass: ML consensus, per-base coverage, candidate regions and the reads over them from drprg_cuda_discover_candidates
against the oracle-side restatement (oracle/discover_py.py) on the same reads."""
import numpy as np
import pytest

import discover_py
import oracle_py as O
from drprg_b200 import lib, sim
from helpers import TOY_PRG, TOY_REFS, panel_sample, small_panel

pytestmark = pytest.mark.gpu


def run(prg, refs, d, o, genome_size, devices=None, **disc):
    gx = lib.Index(prg, 11, 15, devices=devices) if devices else lib.Index(prg, 11, 15, device=0)
    ox = O.Index(prg, 11, 15)
    oo = O.make_opts(illumina=True, genome_size=genome_size)
    go = lib.make_opts(illumina=True, genome_size=genome_size)
    mr = O.MapRun(ox, d, o, oo)
    og = O.Genotype(ox, mr, oo, refs)
    words, woff, lens = lib.pack_reads(d, o, 10)
    gx.retain_hits(True)
    gx.sample_begin(go, 150)
    gx.map_batch(gx.upload(words, woff, lens, total_bases=int(o[-1]), stride_words=10))
    gx.genotype(refs)
    got_loci, got_regions = gx.discover_candidates(**disc)
    kw = {k: v for k, v in disc.items()}
    want_loci, want_regions = discover_py.discover(ox, mr, og, open(prg).read().splitlines(), **kw)
    assert sorted(got_loci) == sorted(want_loci)
    for l in want_loci:
        assert got_loci[l][0] == want_loci[l][0], ("consensus", l)
        assert (got_loci[l][1] == want_loci[l][1]).all(), ("coverage", l)
    assert [r[:5] for r in got_regions] == [r[:5] for r in want_regions]
    for g, w in zip(got_regions, want_regions):
        assert g[5] == w[5], ("reads of region", g[:5])
    return got_loci, got_regions


def test_toy_sample_with_uncovered_stretches():
    """reads tiled from the reference path with two stretches of gid left uncovered: the stretches come back as candidate
    regions with the reads flanking them"""
    d, o = sim.toy_dataset(TOY_PRG, TOY_REFS, depth=40, decoys=2, seed=5)
    refs = [l.strip() for l in open(TOY_REFS) if not l.startswith(">")]
    reads = [d[int(o[i]):int(o[i + 1])].tobytes().decode() for i in range(len(o) - 1)]
    comp = str.maketrans("ACGT", "TGCA")
    hole = refs[0][400:425]
    keep = [r for r in reads if hole not in r and hole not in r.translate(comp)[::-1]]
    assert len(keep) < len(reads)
    dd = np.frombuffer("".join(keep).encode(), np.uint8).copy()
    oo_ = np.arange(len(keep) + 1, dtype=np.uint64) * np.uint64(150)
    loci, regions = run(TOY_PRG, TOY_REFS, dd, oo_, 2000)
    assert len(loci) == 2 and len(regions) >= 1
    assert any(r[0] == 0 and r[5] for r in regions)  # a gid region with supporting reads
    # another threshold / padding combination, and shards sharing the GPU
    run(TOY_PRG, TOY_REFS, dd, oo_, 2000, covg_threshold=8, max_len=60, padding=10)
    run(TOY_PRG, TOY_REFS, dd, oo_, 2000, devices=[0, 0, 0])


def test_panel_sample_low_depth():
    """a 5-locus panel at low depth: many loci have short uncovered runs"""
    p, prg, refs = small_panel()
    d, o, g, pl = panel_sample(p, 12_000, seed=61)
    loci, regions = run(prg, refs, d, o, len(g))
    assert len(loci) >= 3
    run(prg, refs, d, o, len(g), covg_threshold=5, min_len=3, max_len=40)


def test_one_pass_through_the_pandora_mirror(tmp_path):
    """the reference-shaped calls: Pandora.genotype_with(..., retain_hits=True) on a reads FILE (mapped wave by wave), then
    Pandora.discover_candidates() — the same regions and reads as the staged calls on the same reads in memory"""
    from drprg_b200.pandora import Pandora
    p, prg, refs = small_panel()
    d, o, g, pl = panel_sample(p, 12_000, seed=61)
    fq = tmp_path / "reads.fq"
    sim.write_fastq(str(fq), d, o)
    pan = Pandora.from_path()
    pan.genotype_with(prg, refs, fq, tmp_path, ["-t", "4", "-w", "11", "-k", "15", "-c", "10", "-I"], retain_hits=True)
    loci_f, regions_f = pan.discover_candidates()
    # staged reference run (drprg's -g is fixed at the Mtb genome size in genotype_with)
    gx = lib.Index(prg, 11, 15, device=0)
    go = lib.make_opts(illumina=True, genome_size=4411532)
    words, woff, lens = lib.pack_reads(d, o, 10)
    gx.retain_hits(True)
    gx.sample_begin(go, 150)
    gx.map_batch(gx.upload(words, woff, lens, total_bases=int(o[-1]), stride_words=10))
    gx.genotype(refs)
    loci_s, regions_s = gx.discover_candidates()
    assert sorted(loci_f) == sorted(loci_s) and len(regions_f) == len(regions_s) and len(regions_s) > 0
    for l in loci_s:
        assert loci_f[l][0] == loci_s[l][0] and (loci_f[l][1] == loci_s[l][1]).all()
    assert regions_f == regions_s
