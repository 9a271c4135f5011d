"""The SWAR decision flags of k_trace_dp16 (pecaller_b200/csrc/trace_int.cuh) restated with numpy uint32 arithmetic and
checked against direct per-half comparisons: two winners per 32-bit word (s16x2, biased so that every half is in
[0, 2^15)), bit 15 of each half of (a + 0x8000 - b) is set iff a >= b, and an X-decision tie (S + 71 == S0) is marked by
"equal" without the X flag.  Also the accessor's decoding (LaneBandCell) of the six flags into the walker's cell code."""
import numpy as np

H = np.uint32(0x80008000)
K1 = np.uint32(0x00010001)
HX = np.uint32(0x80008000 + 70 * 0x00010001)


def pack(lo, hi):
    return (lo.astype(np.uint32) | (hi.astype(np.uint32) << np.uint32(16))).astype(np.uint32)


def flags(s0, s1, s2):
    """the six flag bits (bit 15 / 31 per half) exactly as the band pass computes them"""
    x01 = pack(np.maximum(s0 & 0xFFFF, s1 & 0xFFFF), np.maximum(s0 >> 16, s1 >> 16))  # __vmaxs2 on non-negative halves
    g = s0 + H - s1
    g2 = s1 + H - s0
    hh = x01 + H - s2
    h2 = s2 + H - x01
    w1 = s1 + HX - s0
    w2 = s2 + HX - s0
    tx1 = (w1 + K1) & ~w1
    tx2 = (w2 + K1) & ~w2
    f0 = ~g & H
    f1 = ~hh & H
    f2 = w1 & H
    f3 = w2 & H
    f4 = ((g & g2) | tx1) & H
    f5 = ((hh & h2) | tx2) & H
    return f0, f1, f2, f3, f4, f5


def decode(f):
    """LaneBandCell::operator(): bits 0-1 argmax, 2 X1, 3 X2, 4 S1==S0, 5 S2==max01, 6 X1 tie, 7 X2 tie"""
    f0, f1, f2, f3, f4, f5 = f
    return (np.where(f1 == 1, 2, f0) | (f2 << 2) | (f3 << 3) | ((f4 & f2) << 4) | ((f5 & f3) << 5) |
            ((f4 & (1 - f2)) << 6) | ((f5 & (1 - f3)) << 7))


def test_swar_flags_equal_direct_comparisons():
    rng = np.random.default_rng(5)
    n = 400_000
    # values around the bias, with many near-ties and S + 71 == S0 cases
    base = rng.integers(300, 6000, size=(2, n))
    d1 = rng.choice([-200, -72, -71, -70, -1, 0, 1, 5, 70, 71, 72], size=(2, n)) + rng.integers(-1, 2, size=(2, n))
    d2 = rng.choice([-200, -72, -71, -70, -1, 0, 1, 5, 70, 71, 72], size=(2, n)) + rng.integers(-1, 2, size=(2, n))
    S0, S1, S2 = base, base + d1, base + d2
    with np.errstate(over="ignore"):
        f = flags(pack(S0[0], S0[1]), pack(S1[0], S1[1]), pack(S2[0], S2[1]))
    for half, sh in ((0, 15), (1, 31)):
        got = [((x >> np.uint32(sh)) & np.uint32(1)).astype(np.int64) for x in f]
        s0, s1, s2 = S0[half], S1[half], S2[half]
        m01 = np.maximum(s0, s1)
        x1_tie, x2_tie = s1 + 71 == s0, s2 + 71 == s0
        assert np.array_equal(got[0], (s1 > s0).astype(np.int64))
        assert np.array_equal(got[1], (s2 > m01).astype(np.int64))
        assert np.array_equal(got[2], (s1 - 1 > s0 - 72).astype(np.int64))       # X1 (1823-1831)
        assert np.array_equal(got[3], (s2 - 1 > s0 - 72).astype(np.int64))       # X2 (1814-1822)
        assert np.array_equal(got[4], ((s1 == s0) | x1_tie).astype(np.int64))
        assert np.array_equal(got[5], ((s2 == m01) | x2_tie).astype(np.int64))
        c = decode(got)
        ak = np.where(s2 > m01, 2, np.where(s1 > s0, 1, 0))
        assert np.array_equal(c & 3, ak)
        assert np.array_equal((c >> 4) & 1, (s1 == s0).astype(np.int64))
        assert np.array_equal((c >> 5) & 1, (s2 == m01).astype(np.int64))
        assert np.array_equal((c >> 6) & 1, x1_tie.astype(np.int64))
        assert np.array_equal((c >> 7) & 1, x2_tie.astype(np.int64))
