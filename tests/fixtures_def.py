"""Seeded fixture definitions shared by the golden-vector generator (tools/make_golden.py),
the CPU tests (oracle vs golden) and the GPU tests (CUDA path vs oracle vs golden).

Each fixture = a genome + one or more read sets + the reference pemapper command line that
was used on it.  The inputs are regenerated from seeds wherever they are needed; only the
reference's OUTPUTS are committed (tests/golden/<fixture>/...).
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pecaller_b200 import synth  # noqa: E402


@dataclass
class RunDef:
    name: str                       # read-set name (file stem)
    paired: bool
    reads1: np.ndarray
    reads2: np.ndarray | None
    min_align: float
    max_dist: int = 500
    min_dist: int = 0
    bisulfite: bool = False


@dataclass
class Fixture:
    name: str
    genome: list
    names: list
    runs: list = field(default_factory=list)
    bisulfite: bool = False         # index built in bisulfite mode (index_genome_whole.c:163, 174-175: C coded as T)


def _edge_reads(genome, length=100):
    rows = []
    for c, g in enumerate(genome):
        L = g.shape[0]
        starts = [0, 1, 5, 14, 15, 16, 17, 40] + [L - length - k for k in (0, 1, 5, 13, 14, 15, 16, 20, 21, 30)]
        rows.append(synth.reads_at(genome, c, starts, length))
        rows.append(synth.reads_at(genome, c, starts, length, reverse=True))
    return np.concatenate(rows)


def fx_cfg1():
    """BASELINE config 1: 1 contig x 1 Mb, 100k single-end 100 bp; error-free set and 1 % subs set."""
    g = synth.random_genome(1, [1_000_000])
    fx = Fixture("cfg1", g, ["chr1"])
    clean = synth.simulate_reads(2, g, 100_000, 100)
    subs = synth.simulate_reads(3, g, 100_000, 100, sub=0.01)
    fx.runs.append(RunDef("clean", False, clean.reads1, None, 0.9))
    fx.runs.append(RunDef("subs", False, subs.reads1, None, 0.9))
    return fx


def fx_pe150():
    """Config-2-shaped, scaled: 1 contig x 2 Mb, 50k pairs 150 bp, 1 % sub, 0.1 % ins, 0.1 % del."""
    g = synth.random_genome(20, [2_000_000])
    fx = Fixture("pe150", g, ["chr20"])
    rs = synth.simulate_reads(21, g, 50_000, 150, paired=True, sub=0.01, ins=0.001, dele=0.001)
    fx.runs.append(RunDef("pairs", True, rs.reads1, rs.reads2, 0.85))
    se = synth.simulate_reads(22, g, 30_000, 150, sub=0.01, ins=0.001, dele=0.001, n_rate=0.002)
    fx.runs.append(RunDef("single", False, se.reads1, None, 0.85))
    return fx


def fx_edge9():
    """9 contigs x ~100 kb: contig-boundary reads (find_chrom / window clamp quirks), N reads, short/odd lengths."""
    lens = [100_000, 90_001, 110_017, 95_500, 100_016, 99_999, 120_000, 80_033, 105_000]
    g = synth.random_genome(9, lens)
    fx = Fixture("edge9", g, [f"ctg{i}" for i in range(9)])
    edge = _edge_reads(g, 100)
    rnd = synth.simulate_reads(10, g, 4000, 100, sub=0.02, ins=0.002, dele=0.002, n_rate=0.01)
    fx.runs.append(RunDef("edge100", False, np.concatenate([edge, rnd.reads1]), None, 0.9))
    pe = synth.simulate_reads(11, g, 3000, 100, paired=True, sub=0.01, ins=0.001, dele=0.001,
                              insert_range=(150, 400))
    fx.runs.append(RunDef("pairs100", True, pe.reads1, pe.reads2, 0.85, max_dist=450, min_dist=50))
    for length in (64, 129, 250):
        r = synth.simulate_reads(12 + length, g, 1500, length, sub=0.01, ins=0.001, dele=0.001)
        fx.runs.append(RunDef(f"len{length}", False, r.reads1, None, 0.9))
    return fx


def _tie_genome():
    """fp64 rounding decides ties.  Three scenarios in one 1 Mb genome; each has a 100-bp
    unit present as two 1-mismatch copies.  A: copies mutated at unit offsets (3, 1) in genome order ->
    reference says unique at the second; B: offsets (1, 3) -> discarded; C: offsets (1, 2) -> discarded."""
    rng = np.random.Generator(np.random.PCG64(77))
    g = synth.ACGT[rng.integers(0, 4, size=1_000_000, dtype=np.uint8)]
    units = synth.ACGT[rng.integers(0, 4, size=(3, 100), dtype=np.uint8)]

    def mut(u, off):
        c = u.copy()
        c[off] = synth.ACGT[(synth._CODE[c[off]] + 1) & 3]
        return c
    places = [(0, 3, 200_000, 1, 700_000), (1, 1, 250_000, 3, 750_000), (2, 1, 300_000, 2, 800_000)]
    for u, o1, p1, o2, p2 in places:
        g[p1:p1 + 100] = mut(units[u], o1)
        g[p2:p2 + 100] = mut(units[u], o2)
    return g, units


def fx_repeat():
    """Config-5-shaped, scaled: 16 contigs x 250 kb repeat genome (2 kb units, 0-2 % divergence) plus the
    fp64 tie scenarios; single-end 150 bp and pairs.  Stresses too_many_spots, the 200-candidate cap,
    order/dedup and rounding-dependent uniqueness."""
    g = synth.repeat_genome(50, [250_000] * 16, unit_len=2000, n_units=40, frac=0.5, max_div=0.02)
    tie, units = _tie_genome()
    genome = g + [tie]
    fx = Fixture("repeat", genome, [f"rep{i}" for i in range(16)] + ["tie"])
    se = synth.simulate_reads(51, genome, 20_000, 150, sub=0.005, ins=0.0005, dele=0.0005)
    fx.runs.append(RunDef("single", False, se.reads1, None, 0.85))
    pe = synth.simulate_reads(52, genome, 8_000, 150, paired=True, sub=0.005, ins=0.0005, dele=0.0005)
    fx.runs.append(RunDef("pairs", True, pe.reads1, pe.reads2, 0.85))
    tie_reads = np.concatenate([units, synth.revcomp_rows(units)])
    fx.runs.append(RunDef("ties", False, tie_reads, None, 0.9))
    return fx


def fx_tiny():
    """Seconds-scale smoke fixture: 1 contig x 60 kb, 2k single + 1k pairs, 100 bp."""
    g = synth.random_genome(5, [60_000])
    fx = Fixture("tiny", g, ["t1"])
    se = synth.simulate_reads(6, g, 2000, 100, sub=0.01, ins=0.001, dele=0.001, n_rate=0.003)
    fx.runs.append(RunDef("single", False, se.reads1, None, 0.9))
    pe = synth.simulate_reads(7, g, 1000, 100, paired=True, sub=0.01, ins=0.001, dele=0.001,
                              insert_range=(150, 400))
    fx.runs.append(RunDef("pairs", True, pe.reads1, pe.reads2, 0.85))
    return fx


def _bisulfite_convert(rng, reads, rate=0.9):
    """Unmethylated C reads as T after bisulfite treatment: convert each C of the read as sequenced with p = rate."""
    out = reads.copy()
    hit = (out == ord("C")) & (rng.random(out.shape) < rate)
    out[hit] = ord("T")
    return out


def fx_bis():
    """Bisulfite mode: 1 contig x 80 kb indexed with C = T, C->T converted reads, IS_BISULFITE = y."""
    g = synth.random_genome(60, [80_000])
    fx = Fixture("bis", g, ["chrB"], bisulfite=True)
    rng = np.random.Generator(np.random.PCG64(61))
    se = synth.simulate_reads(62, g, 2500, 100, sub=0.01, ins=0.0005, dele=0.0005)
    fx.runs.append(RunDef("single", False, _bisulfite_convert(rng, se.reads1), None, 0.85, bisulfite=True))
    pe = synth.simulate_reads(63, g, 1200, 100, paired=True, sub=0.01)
    fx.runs.append(RunDef("pairs", True, _bisulfite_convert(rng, pe.reads1), _bisulfite_convert(rng, pe.reads2), 0.85,
                          bisulfite=True))
    return fx


FIXTURES = {"bis": fx_bis, "tiny": fx_tiny, "cfg1": fx_cfg1, "pe150": fx_pe150, "edge9": fx_edge9, "repeat": fx_repeat}
