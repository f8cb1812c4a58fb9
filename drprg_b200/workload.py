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


class Config3(Config2):
    """configs[2]: the config-2 panel, genome and sample haplotype; 30 M simulated 150 bp Illumina reads (~1000x panel
    depth), read-sharded over the GPUs of the box.  The reads are cut into 240 sub-shards of 125 000 reads with their
    own seeds, so any rank of any world size (1, 2, 4, 8, ...) regenerates exactly its contiguous part: rank r of N owns
    sub-shards [240 r / N, 240 (r + 1) / N).  The sub-shards are generated ON THE GPU with torch (plumbing: the same
    simulator as sim.simulate_reads — uniform start, random strand, 0.2 % substitutions — at ~1 ms per sub-shard
    instead of ~0.3 s in numpy, which would be over a minute for the whole sample)."""

    name = ("config3: synthetic Mtb-scale panel (30 loci, ~4.4k sites, w=11,k=15), 30 M simulated 150 bp Illumina reads "
            "(~1000x panel depth), read-sharded, -I -c 10")
    SUBSHARD = 125_000
    N_SUBSHARDS = 240

    def __init__(self, workdir=None, total_reads=30_000_000):
        super().__init__(workdir)
        if total_reads % self.SUBSHARD:
            raise ValueError("total reads must be a multiple of 125 000")
        self.total_reads = total_reads
        self.n_subshards = total_reads // self.SUBSHARD
        self._genome_dev = {}

    def rank_subshards(self, rank, world):
        return range(self.n_subshards * rank // world, self.n_subshards * (rank + 1) // world)

    def _genome_codes(self, device):
        import torch
        key = str(device)
        if key not in self._genome_dev:
            self._genome_dev[key] = torch.from_numpy(sim._CODE[self.genome].astype(np.uint8)).to(device)
        return self._genome_dev[key]

    def subshard_codes(self, i, device="cuda"):
        """2-bit codes (uint8 tensor [SUBSHARD, 150] on `device`) of sub-shard i"""
        import torch
        gen = torch.Generator(device=device)
        gen.manual_seed(PANEL_SEED + 100_000 + i)
        g = self._genome_codes(device)
        n, L = self.SUBSHARD, READ_LEN
        starts = torch.randint(0, len(self.genome) - L + 1, (n,), generator=gen, device=device)
        codes = g[starts[:, None] + torch.arange(L, device=device)[None, :]]
        strand = torch.rand(n, generator=gen, device=device) < 0.5
        codes = torch.where(strand[:, None], 3 - codes.flip(1), codes)
        err = torch.rand((n, L), generator=gen, device=device) < 0.002
        shift = torch.randint(1, 4, (n, L), generator=gen, device=device, dtype=torch.uint8)
        return torch.where(err, (codes + shift) % 4, codes)

    @staticmethod
    def pack_codes(codes):
        """[n, 150] codes -> [n, STRIDE_WORDS] int32 words in the library's layout (first base in the top bits)"""
        import torch
        n, L = codes.shape
        pad = torch.zeros((n, STRIDE_WORDS * 16), dtype=torch.int64, device=codes.device)
        pad[:, :L] = codes
        shifts = (30 - 2 * torch.arange(16, device=codes.device, dtype=torch.int64))
        w = (pad.view(n, STRIDE_WORDS, 16) << shifts).sum(-1)
        return ((w + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)  # uint32 bit pattern as int32

    def write_fastq(self, path, n_subshards=None, device="cuda"):
        """the first n_subshards sub-shards as one plain 4-line FASTQ (fixed-width ids, constant qualities; 315 bytes per
        read): the file form of the workload for the drop-in call.  Returns the number of reads written."""
        n_sub = self.n_subshards if n_subshards is None else min(n_subshards, self.n_subshards)
        L = READ_LEN
        with open(path, "wb") as f:
            for i in range(n_sub):
                codes = self.subshard_codes(i, device).cpu().numpy()
                n = codes.shape[0]
                rec = np.empty((n, 11 + L + 3 + L + 1), np.uint8)
                ids = np.arange(i * n, (i + 1) * n, dtype=np.int64)
                rec[:, 0], rec[:, 1] = ord("@"), ord("r")
                for d in range(8):
                    rec[:, 2 + d] = ord("0") + (ids // 10 ** (7 - d)) % 10
                rec[:, 10] = 10
                rec[:, 11:11 + L] = sim.BASES[codes]
                rec[:, 11 + L], rec[:, 12 + L], rec[:, 13 + L] = 10, ord("+"), 10
                rec[:, 14 + L:14 + 2 * L] = ord("F")
                rec[:, 14 + 2 * L] = 10
                f.write(rec.tobytes())
        return n_sub * self.SUBSHARD

    def subshard_ascii(self, i, device="cuda"):
        """(data, off) like sim.simulate_reads, on the host (CPU baselines, oracle checks)"""
        codes = self.subshard_codes(i, device).cpu().numpy()
        data = sim.BASES[codes].reshape(-1)
        off = np.arange(codes.shape[0] + 1, dtype=np.uint64) * np.uint64(READ_LEN)
        return data, off
