"""Shared dataset builders for the parity tests (seeded, small enough for the oracle to finish in seconds)."""
import os
import tempfile

import numpy as np

from drprg_b200 import sim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOY_PRG = os.path.join(GOLDEN, "toy.dr.prg")
TOY_REFS = os.path.join(GOLDEN, "toy.genes.fa")

_cache = {}


def reads_from_strings(strs):
    data = np.frombuffer("".join(strs).encode(), np.uint8).copy()
    off = np.zeros(len(strs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in strs])
    return data, off


def small_panel(seed=7, n_loci=5, n_sites=500):
    key = ("panel", seed, n_loci, n_sites)
    if key not in _cache:
        p = sim.make_panel(seed=seed, n_loci=n_loci, n_sites=n_sites, len_lo=900, len_hi=2200,
                           frac=(0.7, 0.12, 0.12, 0.06))
        d = tempfile.mkdtemp(prefix="drprg_panel_")
        prg, fa = p.write(d)
        _cache[key] = (p, prg, fa)
    return _cache[key]


def panel_sample(panel, n_reads, seed=11, genome_size=200_000, alt_frac=0.1, read_len=150, sub_rate=0.002):
    hap = sim.sample_haplotype(panel, seed, alt_frac)
    g, placements = sim.make_genome(panel, [h[0] for h in hap], size=genome_size, seed=seed + 1, min_sep=3000)
    d, o = sim.simulate_reads(g, n_reads, read_len, seed + 2, sub_rate)
    return d, o, g, placements


def long_reads_sample(panel, n_reads, seed=21, genome_size=200_000, mean_len=4000):
    hap = sim.sample_haplotype(panel, seed, 0.1)
    g, placements = sim.make_genome(panel, [h[0] for h in hap], size=genome_size, seed=seed + 1, min_sep=3000)
    d, o = sim.simulate_long_reads(g, n_reads, mean_len=mean_len, seed=seed + 2)
    return d, o, g, placements
