"""CPU: the claim behind diag_score64 (sw_int16.cuh).  When the integer DP (units of 1/36) has a unique last-column
maximum in state 0 and, behind that cell, a diagonal on which S0 beats S1 and S2 by at least 1/36 at every cell
(ITaskResult.flags == 4), the reference's DOUBLE score (smith_waterman_align, pemapper.c:1694-1748, here the pinned
oracle) is the plain left-to-right sum of the bonus values along that diagonal, started from M[i0][0] = 0 or from the
row-0 border: bit for bit, with the same end cell."""
import numpy as np

import oracle_lib as ol

GO, GE, MATCH, MISM = 72, 1, 36, -12


def int_dp(read, win):
    """-> (best, bk, bi, tie, pure): last-column scan of the integer DP and the strict-diagonal flag of its end cell"""
    mm, nn = len(read), len(win)
    NEG = -10**6
    S0 = np.full((nn + 1, mm + 1), NEG); S1 = S0.copy(); S2 = S0.copy()
    E = np.full((nn + 1, mm + 1), 10**6)
    for j in range(1, mm + 1):
        S0[0][j] = S1[0][j] = S2[0][j] = -(GO + (j - 1) * GE)
    S0[:, 0] = 0; S1[:, 0] = 0; S2[:, 0] = -GO
    best, bk, bi, tie = -(GO + (mm - 1) * GE), 0, 0, 0
    for i in range(1, nn + 1):
        for j in range(1, mm + 1):
            S2[i][j] = max(S0[i][j - 1] - GO, S2[i][j - 1] - GE)
            S1[i][j] = max(S0[i - 1][j] - GO, S1[i - 1][j] - GE)
            m = max(S0[i - 1][j - 1], S1[i - 1][j - 1], S2[i - 1][j - 1]) if j > 1 else 0
            S0[i][j] = m + (MATCH if read[j - 1] == win[i - 1] else MISM)
            E[i][j] = min(E[i - 1][j - 1], max(0, S0[i][j] - max(S1[i][j], S2[i][j])))
        for k, v in enumerate((S0[i][mm], S1[i][mm], S2[i][mm])):
            if v > best:
                best, bk, bi, tie = v, k, i, 0
            elif v == best:
                tie = 1
    pure = bk == 0 and bi > 0 and E[bi][mm] >= 1
    return best, bk, bi, tie, pure


def fold(read, win, maxi):
    mm = len(read)
    mism = -1.0 / (3.0 * 1.0)                      # init_bonus_matrices 2011
    j0 = 0 if maxi >= mm else mm - maxi
    i0 = maxi - (mm - j0)
    s = 0.0 if j0 == 0 else -(2.0 + (j0 - 1) * (1.0 / 36.0))   # init_penalty_matrices 2073-2081
    for j in range(j0 + 1, mm + 1):
        s = s + (1.0 if read[j - 1] == win[i0 + (j - j0) - 1] else mism)
    return s


def test_double_score_of_a_strict_diagonal_is_its_running_sum(oracle_built):
    rng = np.random.default_rng(17)
    genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 6000)]
    orc = ol.Oracle([genome])
    used = row0 = row0_used = 0
    for t in range(220):
        mm = int(rng.integers(20, 56))
        nn = mm + int(rng.integers(0, 22))
        at = int(rng.integers(0, len(genome) - nn - 1))
        win = genome[at:at + nn]
        o = int(rng.integers(0, nn - mm + 1))
        read = win[o:o + mm].copy()
        flip = rng.random(mm) < (0.0 if t % 4 == 0 else 0.06)
        read[flip] = rng.integers(0, 4, int(flip.sum()))
        if t % 5 == 4:                              # some reads hang over the window's start: the diagonal leaves through row 0
            head = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 3)]
            read = np.concatenate([head, win[:mm - 3]])
            row0 += 1
        best, bk, bi, tie, pure = int_dp(read, win)
        if not pure or tie:
            continue
        used += 1
        row0_used += bi < mm
        score, (k, i, j) = orc.sw_align(at, nn, read.tobytes())
        assert (k, i) == (0, bi), t
        got = fold(read, win, bi)
        assert np.float64(got).view(np.uint64) == np.float64(score).view(np.uint64), (t, got, score)
        assert abs(score - best / 36.0) < 1e-9
    assert used > 100 and row0_used > 10
    orc.close()
