"""Pin the oracle against every known-answer the reference tree holds for the map path
(SURVEY.md §8c, Appendix A): PRG grammar, site -> VCF record enumeration, genotype likelihoods."""
import glob
import math
import os

import numpy as np
import pytest

import oracle_py as O
from drprg_b200 import sim

FIXTURE_E = {  # expected k-mer depth per fixture, recovered from zero-depth alleles (-2E) / fit
    "ERR4796933.pandora.vcf": 72, "in2.vcf": 238, "in3.vcf": 241, "SRR6824468.vcf": 248,
    "ERR2510634.drprg.vcf": 17, "in.vcf": 96, "in4.vcf": 73,
}
HAND_EDITED = {("in.vcf", "ddn", 627), ("in.vcf", "katG", 1044), ("in4.vcf", "fabG1", 92)}


def vcf_rows(path):
    for line in open(path):
        if line.startswith("#"):
            continue
        f = line.rstrip("\n").split("\t")
        fmt = f[8].split(":")
        val = dict(zip(fmt, f[9].split(":")))
        yield f, val


@pytest.mark.parametrize("name", sorted(FIXTURE_E))
def test_likelihood_known_answers(golden, name):
    """LIKELIHOOD, GT and GT_CONF of every pandora VCF row reproduce from MEAN covgs + GAPS."""
    E = FIXTURE_E[name]
    n_alleles = 0
    for f, v in vcf_rows(os.path.join(golden, name)):
        if (name, f[0], int(f[1])) in HAND_EDITED:
            continue
        mf = [int(x) for x in v["MEAN_FWD_COVG"].split(",")]
        mr = [int(x) for x in v["MEAN_REV_COVG"].split(",")]
        gaps = [float(x) for x in v["GAPS"].split(",")]
        want = [float(x) for x in v["LIKELIHOOD"].split(",")]
        c = [a + b for a, b in zip(mf, mr)]
        got = [O.allele_likelihood(E, ci, sum(c) - ci, g, 0.01) for ci, g in zip(c, gaps)]
        for a, b in zip(got, want):
            # 6 significant digits printed; GAPS is itself rounded to 6 digits and multiplied by E
            assert math.isclose(a, b, rel_tol=2e-5, abs_tol=2e-3), (name, f[0], f[1], got, want)
            n_alleles += 1
        best = int(np.argmax(got))
        if v["GT"] != ".":
            srt = sorted(got, reverse=True)
            if srt[0] - srt[1] > 1e-3:
                assert int(v["GT"]) == best, (name, f[0], f[1])
            assert math.isclose(srt[0] - srt[1], float(v["GT_CONF"]), rel_tol=2e-4, abs_tol=5e-3)
    assert n_alleles >= 12


def test_zero_depth_is_minus_2E():
    assert O.allele_likelihood(72, 0, 0, 1.0, 0.01) == -144.0


def test_toy_prg_grammar(golden):
    ix = O.Index(os.path.join(golden, "toy.dr.prg"), 11, 15)
    assert ix.names == ["gid", "pncA"]
    text = open(os.path.join(golden, "toy.dr.prg")).read().splitlines()
    for li, body in ((0, text[1]), (1, text[3])):
        lg = ix.local_graph(li)
        toks = body.split(" ")
        dna = [t for t in toks if not t.isdigit()]
        # one LocalNode per sequence token (empty tokens included), ids in order of appearance
        assert len(lg["start"]) == len(dna)
        assert [int(x) for x in lg["len"]] == [len(t) for t in dna]
        # interval coordinates are PRG-string character offsets
        for s, l, t in zip(lg["start"], lg["len"], dna):
            assert body[int(s):int(s) + int(l)] == t
    # gid: 15 sites (markers 5..34); pncA: 9 sites incl. the nested 13 > 15
    lg = ix.local_graph(0)
    assert int((lg["n_out"] > 1).sum()) == 15
    lg = ix.local_graph(1)
    assert int((lg["n_out"] > 1).sum()) == 9


# (locus, POS, REF, ALTs, VC, GRAPHTYPE); rows marked True appear verbatim
# in real pandora output (in.vcf / SRR6824468.vcf) for the production index sharing these sites.
TOY_SITES = [
    ("gid", 117, "C", "T", "SNP", "SIMPLE"),
    ("gid", 160, "GC", "G", "INDEL", "SIMPLE"),
    ("gid", 269, "G", "A", "SNP", "SIMPLE"),
    ("gid", 303, "T", "G", "SNP", "SIMPLE"),
    ("gid", 330, "TGCCATTGGCGATAGCGCG", "GGCTACGTCACGCACATTT", "PH_SNPs", "SIMPLE"),
    ("gid", 386, "C", "A", "SNP", "SIMPLE"),
    ("gid", 505, "G", "T", "SNP", "NESTED"),
    ("gid", 505, "GTCACGG", "TTGGGCGGCAGCGACGCT", "COMPLEX", "NESTED"),
    ("gid", 608, "G", "A", "SNP", "SIMPLE"),
    ("gid", 813, "G", "C", "SNP", "SIMPLE"),
    ("pncA", 180, "T", "C", "SNP", "SIMPLE"),
    ("pncA", 269, "CACT", "CACC,GACT,TACT", "PH_SNPs", "SIMPLE"),
    ("pncA", 292, "TTCC", "TATCT", "COMPLEX", "SIMPLE"),
    ("pncA", 302, "TGGCC", "GGGCC,TGGCCACCGCATT", "PH_SNPs", "NESTED"),
    ("pncA", 381, "T", "C", "SNP", "SIMPLE"),
    ("pncA", 489, "T", "TG", "INDEL", "SIMPLE"),
    ("pncA", 760, "GG", "AG", "PH_SNPs", "SIMPLE"),
]


@pytest.fixture(scope="module")
def toy_run(golden):
    prg, fa = os.path.join(golden, "toy.dr.prg"), os.path.join(golden, "toy.genes.fa")
    ix = O.Index(prg, 11, 15)
    d, o = sim.toy_dataset(prg, fa, depth=60, decoys=2, seed=1)
    opts = O.make_opts(illumina=True, genome_size=2000)
    mr = O.MapRun(ix, d, o, opts)
    gt = O.Genotype(ix, mr, opts, fa)
    return ix, mr, gt


def test_toy_site_enumeration(toy_run, golden):
    ix, mr, gt = toy_run
    rows = {}
    for line in gt.vcf().splitlines():
        if line.startswith("#"):
            continue
        f = line.split("\t")
        info = dict(x.split("=") for x in f[7].split(";"))
        rows[(f[0], int(f[1]), f[3])] = (f[4], info["VC"], info["GRAPHTYPE"])
    for chrom, pos, ref, alts, vc, gtype in TOY_SITES:
        assert (chrom, pos, ref) in rows, (chrom, pos, ref, sorted(k for k in rows if k[0] == chrom))
        assert rows[(chrom, pos, ref)] == (alts, vc, gtype), (chrom, pos, rows[(chrom, pos, ref)])
    # every record's REF equals the --vcf-refs slice (src/consequence.rs:100-113 hard-fails otherwise)
    refs = {}
    for line in open(os.path.join(golden, "toy.genes.fa")):
        if line.startswith(">"):
            name = line[1:].strip()
        else:
            refs[name] = line.strip()
    for (chrom, pos, ref) in rows:
        assert refs[chrom][pos - 1:pos - 1 + len(ref)] == ref
    assert len(rows) == 15 + 9 - 1  # pncA sites 13 and 15 merge into one multi-allelic record


def test_toy_genotypes_are_reference(toy_run):
    """Reads were tiled from the --vcf-refs (top) path, so every covered site must call GT=0."""
    ix, mr, gt = toy_run
    r = gt.records()
    assert (mr.locus_reads() > 0).all()
    covered = (r["gt_conf"] > 0)
    assert covered.sum() >= 20
    assert (r["gt"][covered] == 0).all()


def test_vcf_schema_matches_fixture_header(toy_run, golden):
    ix, mr, gt = toy_run
    ours = [l for l in gt.vcf().splitlines() if l.startswith("##") and not l.startswith(("##fileDate", "##contig"))]
    theirs = [l for l in open(os.path.join(golden, "ERR4796933.pandora.vcf")).read().splitlines()
              if l.startswith("##") and not l.startswith(("##fileDate", "##contig"))]
    assert ours == theirs
    fmt = [l.split("\t")[8] for l in gt.vcf().splitlines() if not l.startswith("#")][0]
    theirs_fmt = [l.split("\t")[8] for l in open(os.path.join(golden, "in.vcf")) if not l.startswith("#")][0]
    assert fmt == theirs_fmt
