#!/usr/bin/env python
"""CPU model of the seed stage's single-chain shortcut (seed_rbi.cuh, phases M_A / M_B of rbi_map_read_mate).

Test infrastructure: the model restates find_matches (pemapper.c:2189-2288) and initial_map's two calls of it
(1655-1659) over the full 49-k-mer segment lists (1594-1637), then runs the shortcut, which sees only
  * rotation 0 of every segment (the exact k-mer and the 12 substitutions of its last four bases),
  * all four rotations of the first k = nseg - F0 + 2 segments and of the segments rotation 0 left without an entry
    near the chain's diagonal,
and claims the outcome of find_matches for the strand from that alone whenever its conditions hold.  main() checks on
adversarial genomes (tandem repeats, diverged copies, crowded k-mers, low complexity) that every claim equals the full
computation.  tests/test_seed_shortcut_model.py runs it on the CPU; the GPU parity tests cover the kernel itself.
"""
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


def try_single_chain(offsets, segs, min_match, tot_before, max_hits, stats):
    """The shortcut.  segs[s] = (full, [(pos, rot)], crowded[4]).  -> None (no claim) or (new min_match, hit (pos, off))."""
    nseg = len(segs)
    mo = MAX_OFF
    # rotation 0 of every segment; a crowded marker met there makes no claim
    if any(sg[2][0] for sg in segs):
        return None
    e0 = [(p, s) for s, sg in enumerate(segs) for p, rot in sg[1] if rot == 0]
    stats["buckets"] += nseg
    # best anchor over rotation 0 alone (found as in 2230-2249)
    best_f, best_d = 0, None
    for p, s in e0:
        d = p - offsets[s]
        later = {sq for q, sq in e0 if sq > s and abs((q - offsets[sq]) - d) < mo}
        f = 1 + len(later)
        if f > best_f:
            best_f, best_d = f, d
    f0 = best_f
    k = nseg - f0 + 2
    if f0 <= min_match or k > nseg // 2:
        return None
    d = best_d
    covered = {s for p, s in e0 if abs((p - offsets[s]) - d) < mo}
    fetch = set(range(k)) | {s for s in range(k, nseg) if s not in covered}
    for s in range(nseg):
        if s in fetch:
            if any(segs[s][2]):
                return None              # a marker came along: the full path sorts it out
        elif any(segs[s][2][1:]):
            return None                  # directory flag of an unread bucket: it may hold a marker for this segment
    stats["buckets"] += 3 * len(fetch)
    ents = [(p, s) for p, s in e0 if s not in fetch] + [(p, s) for s in fetch for p, _ in segs[s][1]]
    S = [(p, s) for p, s in ents if s < k]
    if tot_before + len(S) >= max_hits:
        return None
    # condition C: inside the first k segments only entries exactly on d may pair up across segments
    for p, s in S:
        de = p - offsets[s]
        if de == d:
            continue
        for q, sq in S:
            if sq != s and abs((q - offsets[sq]) - de) < 2 * mo - 1:
                return None
    on_d = sorted(s for p, s in ents if p - offsets[s] == d)
    s_first = on_d[0]
    cov2 = {s for p, s in ents if abs((p - offsets[s]) - d) < mo}
    f_max = 1 + len([s for s in cov2 if s > s_first])
    if f_max < f0:
        return None                      # (cannot happen; keeps the claim honest)
    stats["claims"] += 1
    return f_max, (d + offsets[s_first], offsets[s_first])


def initial_map(ix, read, max_hits, shortcut, stats):
    """-> hits [(pos - off clipped at 0, orient)] as initial_map returns them (1661-1669)"""
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    length = len(read)
    n_count = read.count(b"N")
    if length < 16 or n_count >= 1 + length // 10:
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
        claim = try_single_chain(offsets, segs, min_match, len(hits), max_hits, stats) if shortcut else None
        if claim is not None:
            min_match = claim[0]
            del hits[:]
            hits.append((claim[1][0], claim[1][1], orient))
        else:
            if shortcut:
                stats["buckets"] += 4 * len(segs)
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


def main(n_reads=400, seed=5, max_hits_list=(200, 12), noise=0.0, n_random=300_000):
    rng = np.random.default_rng(seed)
    genome = make_genome(rng, n_random)
    ix = Index(genome, noise)
    reads = make_reads(rng, genome, n_reads)
    report = {}
    for max_hits in max_hits_list:
        st_full = dict(strands=0, buckets=0, claims=0)
        st_cut = dict(strands=0, buckets=0, claims=0)
        bad = 0
        for r in reads:
            a = initial_map(ix, r, max_hits, False, st_full)
            b = initial_map(ix, r, max_hits, True, st_cut)
            if a != b:
                bad += 1
                print("MISMATCH", r, a, b, file=sys.stderr)
        report[max_hits] = dict(reads=len(reads), mismatches=bad, strands=st_cut["strands"], claims=st_cut["claims"],
                                buckets_read=st_cut["buckets"], buckets_full=4 * 10 * st_cut["strands"])
    return report


if __name__ == "__main__":
    import json
    rep = main(int(sys.argv[1]) if len(sys.argv) > 1 else 400, noise=float(sys.argv[2]) if len(sys.argv) > 2 else 0.0,
               n_random=int(sys.argv[3]) if len(sys.argv) > 3 else 300_000)
    print(json.dumps(rep, indent=1))
    sys.exit(1 if any(v["mismatches"] for v in rep.values()) else 0)
