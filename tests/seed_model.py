"""CPU model of the seed stage's list semantics (test infrastructure): the 49-k-mer segment lists of a read
(pemapper.c:1594-1637, too_many_spots veto included), find_matches (2189-2288) and initial_map's two calls of it
(1655-1659), restated in plain Python over a small k-mer dictionary, plus the DEAD-STRAND rule of the seed kernel
(seed_rbi.cuh, close_pair_exists): with the running min_match at F, the first k = nseg - F + 2 segments decide whether a
strand can change the hit list at all.  tests/test_dead_strand_rule.py checks the rule against the full computation on
adversarial genomes (tandem repeats, diverged copies, crowded k-mers, low complexity, synthetic chance hits)."""
import sys

import numpy as np

MAX_OFF = 12            # maxim(2, idepth - 4), idepth = 16 (2196)
TOO_MANY = 100          # too_many_spots (1602)


def codes_of(seq):
    """2-bit codes (A C G T -> 0 1 2 3; anything else -> 0, cv[] 2379-2383) of every 16-mer of seq."""
    lut = np.zeros(256, dtype=np.uint64)
    for ch, v in zip(b"ACGT", range(4)):
        lut[ch] = v
    c = lut[seq]
    n = len(seq) - 15
    out = np.zeros(n, dtype=np.uint64)
    for i in range(16):
        out = (out << np.uint64(2)) | c[i:i + n]
    return out


def _mix(x):
    x = (x ^ (x >> 31)) * 0x7FB5D329728EA185 & 0xFFFFFFFFFFFFFFFF
    x = (x ^ (x >> 27)) * 0x81DADEF4BC2DD44D & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 33)


class Index:
    """k-mer -> positions.  noise > 0 adds, for that fraction of ALL 2^32 codes, one made-up position (a function of
    the code): the chance hits a 3 Gb genome gives every segment (~35 per 49 k-mers) on a genome small enough for
    chance pairs and chance chains to be frequent, which is what the shortcut's conditions have to survive."""

    def __init__(self, genome, noise=0.0):
        self.genome = genome
        self.noise = noise
        codes = codes_of(genome)
        self.order = np.argsort(codes, kind="stable").astype(np.uint32)
        self.sorted_codes = codes[self.order]

    def positions(self, code):
        lo = np.searchsorted(self.sorted_codes, code, "left")
        hi = np.searchsorted(self.sorted_codes, code, "right")
        p = self.order[lo:hi]
        if self.noise > 0 and len(p) < TOO_MANY - 1:
            h = _mix(int(code) + 0x9E3779B97F4A7C15)
            if (h & 0xFFFF) < self.noise * 65536:
                extra = np.uint32((h >> 16) % (len(self.genome) - 16))
                if extra not in p:
                    p = np.append(p, extra)
        return p


def variants(code):
    """The 49 k-mers of a segment: exact, then one substitution per base; -> [(code, rotation)].
    Rotation g holds the variants whose changed base lies in byte g of the code (bases 12-4g .. 15-4g)."""
    out = [(code, 0)]
    for b in range(16):
        sh = 2 * (15 - b)
        cur = (code >> sh) & 3
        for v in range(4):
            if v != cur:
                out.append(((code & ~(3 << sh)) | (v << sh), (15 - b) // 4))
    return out


def segment_lists(ix, read_codes, length):
    """-> offsets, per segment: (sorted positions [full list], [(pos, rot)], crowded per rotation)"""
    total_cuts = length // 16
    if length % 16 == 0:
        total_cuts -= 1
    offsets = [16 * s for s in range(total_cuts)] + [length - 16]
    segs = []
    for off in offsets:
        code = int(read_codes[off])
        ent, crowded = [], [False] * 4
        for vc, rot in variants(code):
            p = ix.positions(np.uint64(vc))
            if len(p) >= TOO_MANY:
                crowded[rot] = True
            else:
                ent.extend((int(x), rot) for x in p)
        full = [] if any(crowded) else sorted(x for x, _ in ent)
        segs.append((full, ent, crowded))
    return offsets, segs


def find_matches(lists, offsets, min_match, hits, orient, max_hits):
    """pemapper.c:2189-2288; lists[s] sorted; hits = [(pos, off, orient)] is modified in place; -> min_match"""
    max_depth = len(lists) - 1
    if min(len(l) for l in lists) > max_hits:
        del hits[:]
        return min_match
    loop = 0
    while loop <= 1 + max_depth - min_match:
        for p in lists[loop]:
            found = 1
            for j in range(loop + 1, max_depth + 1):
                for q in lists[j]:
                    if abs((p - q) - (offsets[loop] - offsets[j])) < MAX_OFF:
                        found += 1
                        break
            if found > min_match:
                min_match = found
                del hits[:]
                hits.append((p, offsets[loop], orient))
            elif found == min_match:
                if len(hits) < max_hits:
                    if all(h[0] - h[1] != p - offsets[loop] for h in hits):
                        hits.append((p, offsets[loop], orient))
                else:
                    return min_match
        loop += 1
    return min_match


def strand_is_dead(offsets, segs, min_match, max_hits):
    """The kernel's rule.  -> True when it declares the strand dead (then find_matches must leave hits and min_match
    alone), False when it makes no statement.  segs[s] = (full list, ...)."""
    nseg = len(segs)
    k = nseg - min_match + 2
    if k > nseg // 2:
        return False                      # the probe is only used when it halves the strand
    first = [segs[s][0] for s in range(k)]
    if min(len(l) for l in first) > max_hits:
        return False                      # the min_spots rule (2200-2207) could still fire
    ent = [(p - offsets[s], s) for s in range(k) for p in first[s]]
    for d, s in ent:
        for d2, s2 in ent:
            if s2 != s and abs(d2 - d) < 2 * MAX_OFF - 1:
                return False              # a close pair: some anchor may reach min_match
    return True


def initial_map(ix, read, max_hits, use_rule, stats):
    """-> hits [(pos - off clipped at 0, orient)] as initial_map returns them (1661-1669)"""
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    length = len(read)
    if length < 16 or read.count(b"N") >= 1 + length // 10:
        return []
    fwd = np.frombuffer(read, dtype=np.uint8)
    rev = np.frombuffer(read.translate(comp)[::-1], dtype=np.uint8)
    hits = []
    total_cuts = length // 16 - (1 if length % 16 == 0 else 0)
    min_match = max(1, total_cuts)
    if total_cuts > 4:
        min_match = (4 * total_cuts) // 5
    min_match = min(min_match, 4)
    for orient, seq in ((0, fwd), (1, rev)):
        if orient == 1 and len(hits) >= max_hits:
            break
        offsets, segs = segment_lists(ix, codes_of(seq), length)
        stats["strands"] += 1
        if use_rule and strand_is_dead(offsets, segs, min_match, max_hits):
            stats["dead"] += 1
            continue
        min_match = find_matches([sg[0] for sg in segs], offsets, min_match, hits, orient, max_hits)
    return [(max(0, p - o), orr) for p, o, orr in hits]


def make_genome(rng, n_random=300_000):
    parts = [rng.integers(0, 4, n_random)]
    unit = rng.integers(0, 4, 2000)
    for _ in range(12):                                   # diverged copies of one unit (cfg5-like)
        u = unit.copy()
        m = rng.random(2000) < rng.uniform(0, 0.02)
        u[m] = rng.integers(0, 4, int(m.sum()))
        parts += [u, rng.integers(0, 4, 500)]
    for period in (1, 2, 3, 7, 16, 17, 40):               # tandem repeats
        parts += [np.tile(rng.integers(0, 4, period), 600 // period + 1), rng.integers(0, 4, 300)]
    hot = rng.integers(0, 4, 40)                          # a 40-mer present 130 times: its 16-mers are crowded
    for _ in range(130):
        parts += [hot, rng.integers(0, 4, 30)]
    low = rng.choice([0, 3], 3000)                        # low complexity
    parts += [low, rng.integers(0, 4, 20_000)]
    g = np.concatenate(parts)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[g]


def make_reads(rng, genome, n, length=150):
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    out = []
    G = len(genome)
    for i in range(n):
        L = length if i % 7 else int(rng.integers(40, 200))
        at = int(rng.integers(0, G - L - 10))
        r = bytearray(genome[at:at + L + 8].tobytes())
        j = 0
        while j < len(r):                                  # 1 % substitutions, 0.2 % indels
            u = rng.random()
            if u < 0.01:
                r[j] = b"ACGT"[int(rng.integers(0, 4))]
            elif u < 0.011:
                del r[j]
                continue
            elif u < 0.012:
                r.insert(j, b"ACGT"[int(rng.integers(0, 4))])
                j += 1
            j += 1
        r = bytes(r[:L])
        if len(r) < 16:
            continue
        if i % 29 == 0:
            r = r[:20] + b"N" + r[21:]
        if rng.random() < 0.5:
            r = r.translate(comp)[::-1]
        out.append(r)
    return out


def main(n_reads=300, seed=5, max_hits_list=(200, 12), noise=0.0, n_random=300_000):
    rng = np.random.default_rng(seed)
    genome = make_genome(rng, n_random)
    ix = Index(genome, noise)
    reads = make_reads(rng, genome, n_reads)
    report = {}
    for max_hits in max_hits_list:
        st_full = dict(strands=0, dead=0)
        st_rule = dict(strands=0, dead=0)
        bad = 0
        for r in reads:
            if initial_map(ix, r, max_hits, False, st_full) != initial_map(ix, r, max_hits, True, st_rule):
                bad += 1
        report[max_hits] = dict(reads=len(reads), mismatches=bad, strands=st_rule["strands"], dead=st_rule["dead"])
    return report
