"""CPU: the arithmetic of the rotated bucket index (seed_rbi.cuh), restated in Python.
 * rbi_bucket / rbi_tag: the 48 one-substitution neighbours of a 16-mer split into four groups of 12 by the byte of the
   code they change; the neighbours of group g share the exact k-mer's bucket of rotation g and differ from its tag in
   exactly one 2-bit field, so four bucket reads + a tag filter return the 49 lists of a segment (fill_mers 1969-2003).
 * the SWAR tag filter of the gather loop: for four packed tag bytes, bit 7 of byte k is set iff tag k differs from the
   segment's tag in at most one 2-bit field, the exact tag being kept in rotation 0 only - checked exhaustively over all
   256 x 256 (tag, segment tag) pairs in every byte lane."""
import numpy as np


def rbi_tag(code, g):
    return (code >> (8 * g)) & 255


def rbi_bucket(code, g):
    lo = code & ((1 << (8 * g)) - 1) if g else 0
    hi = 0 if g == 3 else code >> (8 * g + 8)
    return (hi << (8 * g)) | lo


def swar_hits(T, etagx, keep_exact):
    M = 0xFFFFFFFF
    X = T ^ etagx
    D = (X | (X >> 1)) & 0x55555555
    Z = D & (((D | 0x80808080) - 0x01010101) & M)
    hit = (((Z + 0x7F7F7F7F) & M) & 0x80808080) ^ 0x80808080
    hit &= (((D + 0x7F7F7F7F) & M) & 0x80808080) | keep_exact
    return hit


def fields_differing(a, b):
    x = a ^ b
    return sum(1 for f in range(4) if (x >> (2 * f)) & 3)


def test_neighbours_share_the_bucket_of_their_rotation():
    rng = np.random.default_rng(9)
    for code in [0, 0xFFFFFFFF, 0x12345678] + [int(x) for x in rng.integers(0, 1 << 32, 40, dtype=np.uint64)]:
        seen = {g: set() for g in range(4)}
        for f in range(16):
            for d in (1, 2, 3):
                v = code ^ (d << (2 * f))
                g = f // 4                                   # the byte of the code that changed
                assert rbi_bucket(v, g) == rbi_bucket(code, g)
                assert fields_differing(rbi_tag(v, g), rbi_tag(code, g)) == 1
                for other in range(4):                       # ... and in no other rotation's bucket
                    if other != g:
                        assert rbi_bucket(v, other) != rbi_bucket(code, other)
                seen[g].add(v)
        assert all(len(s) == 12 for s in seen.values())
        # bucket and tag together are the code again
        for g in range(4):
            b, t = rbi_bucket(code, g), rbi_tag(code, g)
            lo = b & ((1 << (8 * g)) - 1)
            hi = b >> (8 * g)
            assert (hi << (8 * g + 8)) | (t << (8 * g)) | lo == code


def test_swar_filter_is_the_one_field_rule_exhaustively():
    tags = np.arange(256, dtype=np.uint64)
    for lane in range(4):
        for et in range(256):
            etagx = et * 0x01010101
            for keep in (0x80808080, 0):
                # all 256 tags in byte `lane`, the other lanes hold a tag two fields away (never qualifies)
                filler = et ^ 0x0F
                base = sum(filler << (8 * k) for k in range(4) if k != lane)
                for t in range(256):
                    hit = swar_hits(base | (t << (8 * lane)), etagx, keep)
                    nd = fields_differing(t, et)
                    want = nd == 1 or (nd == 0 and keep != 0)
                    assert bool(hit & (0x80 << (8 * lane))) == want, (lane, et, t, keep)
                    assert hit & ~(0x80 << (8 * lane)) == 0
    assert len(tags) == 256
