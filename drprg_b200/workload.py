"""The benchmark workloads of BASELINE.json / SURVEY.md §8d, built from fixed seeds."""
from __future__ import annotations

import os
import tempfile

import numpy as np

from . import sim

PANEL_SEED = 20231017
GENOME_SIZE = 4_411_532
READ_LEN = 150
STRIDE_WORDS = 10  # 150 bp -> 38 B of 2-bit bases, stored in 40 B


class Config2:
    """configs[1]: synthetic M. tuberculosis-scale panel PRG (30 loci, ~4.4k sites), reads simulated from a
    4.41 Mb genome carrying a sample haplotype (alt allele at 10 % of sites), 150 bp, 0.2 % substitutions."""

    name = "config2: synthetic Mtb-scale panel (30 loci, ~4.4k sites, w=11,k=15), simulated 150 bp Illumina reads, -I -c 10"

    def __init__(self, workdir=None):
        self.panel = sim.make_panel(seed=PANEL_SEED)
        self.workdir = workdir or tempfile.mkdtemp(prefix="drprg_cfg2_")
        self.prg_path, self.refs_path = self.panel.write(self.workdir)
        hap = sim.sample_haplotype(self.panel, PANEL_SEED + 1, 0.10)
        self.genome, self.placements = sim.make_genome(self.panel, [h[0] for h in hap], size=GENOME_SIZE, seed=PANEL_SEED + 5)
        self.w, self.k = 11, 15

    def reads(self, n_reads, shard=0):
        """ASCII reads of shard `shard` (seed + 2 + shard)."""
        return sim.simulate_reads(self.genome, n_reads, READ_LEN, seed=PANEL_SEED + 2 + shard, sub_rate=0.002)
