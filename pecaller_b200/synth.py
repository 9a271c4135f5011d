"""Seeded synthetic genomes and reads for the PEMapper hot path.

Everything here is deterministic in (seed, arguments): numpy's PCG64 integer
streams are platform independent, so the container that generates the golden
vectors and the GPU box that replays them see the same bytes.

Genomes are lists of upper-case ASCII contigs (uint8 arrays).  Reads are
fixed-length upper-case ASCII rows of a (n, length) uint8 matrix, which is
also the layout the C-ABI's contiguous batch entry point takes.
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b
_CODE = np.zeros(256, dtype=np.uint8)
_CODE[ord("C")] = 1
_CODE[ord("G")] = 2
_CODE[ord("T")] = 3


def revcomp_rows(rows: np.ndarray) -> np.ndarray:
    """Reverse-complement every row of an (n, L) ASCII matrix (ACGT only; others -> N)."""
    return _COMP[rows[:, ::-1]]


def random_genome(seed: int, contig_lens) -> list[np.ndarray]:
    """i.i.d. uniform ACGT contigs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [ACGT[rng.integers(0, 4, size=int(n), dtype=np.uint8)] for n in contig_lens]


def repeat_genome(seed: int, contig_lens, unit_len=2000, n_units=200, frac=0.5, max_div=0.02) -> list[np.ndarray]:
    """High-repeat genome (config 5): `frac` of every contig is covered by copies of a
    library of `n_units` units of `unit_len` bp, each copy diverged by U[0,max_div] substitutions."""
    rng = np.random.Generator(np.random.PCG64(seed))
    units = ACGT[rng.integers(0, 4, size=(n_units, unit_len), dtype=np.uint8)]
    out = []
    for n in contig_lens:
        n = int(n)
        g = ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
        n_copies = int(frac * n / unit_len)
        if n_copies:
            slots = rng.permutation(n // unit_len)[:n_copies]
            which = rng.integers(0, n_units, size=n_copies)
            div = rng.random(n_copies) * max_div
            for s, w, d in zip(slots, which, div):
                cp = units[w].copy()
                k = rng.binomial(unit_len, d)
                if k:
                    p = rng.integers(0, unit_len, size=k)
                    cp[p] = ACGT[(_CODE[cp[p]] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3]
                g[s * unit_len:(s + 1) * unit_len] = cp
        out.append(g)
    return out


@dataclass
class ReadSet:
    reads1: np.ndarray            # (n, L) uint8 ASCII
    reads2: np.ndarray | None     # (n, L) uint8 ASCII or None (single-end)
    contig: np.ndarray            # (n,) int32 true contig of mate 1's fragment
    start: np.ndarray             # (n,) int64 0-based contig offset of the fragment's left end
    reverse: np.ndarray           # (n,) bool: fragment taken from the reverse strand


def _extract(rng, genome_cat, gstart, length, sub, ins, dele, chunk=500_000):
    """Walk `length` read bases from concatenated-genome offsets `gstart` with the error model.
    Returns an (n, length) ASCII matrix in forward-genome orientation."""
    n = gstart.shape[0]
    out = np.empty((n, length), dtype=np.uint8)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = hi - lo
        step = np.ones((m, length), dtype=np.int16)
        inserted = np.zeros((m, length), dtype=bool)
        if dele > 0:
            k = rng.binomial(m * length, dele)
            if k:
                r = rng.integers(0, m, size=k)
                p = rng.integers(1, length, size=k)
                d = rng.integers(1, 4, size=k)
                step[r, p] += d.astype(np.int16)
        if ins > 0:
            k = rng.binomial(m * length, ins)
            if k:
                r = rng.integers(0, m, size=k)
                p = rng.integers(1, length - 3, size=k)
                d = rng.integers(1, 4, size=k)
                for off in range(3):
                    sel = d > off
                    inserted[r[sel], p[sel] + off] = True
                step[inserted] = 0
        step[:, 0] = 0
        idx = gstart[lo:hi, None] + np.cumsum(step, axis=1, dtype=np.int64)
        np.minimum(idx, genome_cat.shape[0] - 1, out=idx)
        rows = genome_cat[idx]
        n_ins = int(inserted.sum())
        if n_ins:
            rows[inserted] = ACGT[rng.integers(0, 4, size=n_ins, dtype=np.uint8)]
        if sub > 0:
            k = rng.binomial(m * length, sub)
            if k:
                r = rng.integers(0, m, size=k)
                p = rng.integers(0, length, size=k)
                rows[r, p] = ACGT[(_CODE[rows[r, p]] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3]
        out[lo:hi] = rows
    return out


def simulate_reads(seed: int, genome: list[np.ndarray], n: int, length: int, paired: bool = False,
                   sub: float = 0.0, ins: float = 0.0, dele: float = 0.0,
                   insert_range=(250, 450), n_rate: float = 0.0) -> ReadSet:
    """Sample `n` reads (or pairs) of fixed `length`.

    Single-end: the read is the fragment's first `length` bases; strand with p=0.5.
    Paired: fragment length U[insert_range]; mate 1 = left end forward, mate 2 = reverse
    complement of the right end; with p=0.5 the mates swap (fragment from the reverse strand).
    sub/ins/dele are per-base rates; insertions and deletions are 1-3 bp (SURVEY.md §8d cfg 2).
    n_rate replaces bases by 'N' after everything else.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([g.shape[0] for g in genome], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    cat = np.concatenate(genome)
    margin = 16  # room for deletions
    frag = rng.integers(insert_range[0], insert_range[1] + 1, size=n) if paired else np.full(n, length)
    frag = np.maximum(frag, length)
    usable = lens - (int(frag.max()) + margin)
    if (usable <= 0).any():
        raise ValueError("contig shorter than fragment + margin")
    contig = rng.choice(len(genome), size=n, p=usable / usable.sum()).astype(np.int32)
    start = (rng.random(n) * (lens[contig] - frag - margin)).astype(np.int64)
    reverse = rng.random(n) < 0.5
    g0 = offs[contig] + start
    left = _extract(rng, cat, g0, length, sub, ins, dele)
    if paired:
        right = revcomp_rows(_extract(rng, cat, g0 + frag - length, length, sub, ins, dele))
        r1 = np.where(reverse[:, None], right, left)
        r2 = np.where(reverse[:, None], left, right)
    else:
        r1 = np.where(reverse[:, None], revcomp_rows(left), left)
        r2 = None
    if n_rate > 0:
        for r in (r1, r2):
            if r is not None:
                k = rng.binomial(r.size, n_rate)
                r.reshape(-1)[rng.integers(0, r.size, size=k)] = ord("N")
    return ReadSet(np.ascontiguousarray(r1), None if r2 is None else np.ascontiguousarray(r2),
                   contig, start, reverse)


def reads_at(genome: list[np.ndarray], contig: int, starts, length: int, reverse=False) -> np.ndarray:
    """Error-free reads at explicit contig offsets (edge fixture, SURVEY.md §8d 'edge')."""
    g = genome[contig]
    rows = np.stack([g[s:s + length] for s in starts])
    return revcomp_rows(rows) if reverse else rows


def write_fasta(path: str, genome: list[np.ndarray], names=None, width: int = 60) -> None:
    with open(path, "wb") as f:
        for i, g in enumerate(genome):
            name = names[i] if names else f"chr{i + 1}"
            f.write(b">" + name.encode() + b"\n")
            pad = (-g.shape[0]) % width
            body = np.concatenate([g, np.zeros(pad, np.uint8)]).reshape(-1, width)
            lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).reshape(-1)
            data = lines.tobytes().replace(b"\x00", b"")
            if not data.endswith(b"\n"):
                data += b"\n"
            f.write(data)


def write_fastq(path: str, reads: np.ndarray, prefix: str = "r") -> None:
    """FASTQ with a constant quality string; plain or .gz by extension."""
    n, length = reads.shape
    qual = b"I" * length
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "wb") as f:
        buf = []
        for i in range(n):
            buf.append(b"@%s%d\n%s\n+\n%s\n" % (prefix.encode(), i, reads[i].tobytes(), qual))
            if len(buf) >= 65536:
                f.write(b"".join(buf))
                buf = []
        f.write(b"".join(buf))


def human_like_contig_lens(total: int, n: int = 24) -> list[int]:
    """Contig sizes proportional to GRCh38 chr1-22,X,Y scaled to `total` bp (config 3)."""
    hs = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28,
          114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23][:n]
    s = sum(hs)
    lens = [int(total * h / s) for h in hs]
    lens[0] += total - sum(lens)
    return lens


def ensure_dir(path: str) -> str:
    os.makedirs(path, exist_ok=True)
    return path
