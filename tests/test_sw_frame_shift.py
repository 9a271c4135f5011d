"""CPU: the frame k_sw_i16 keeps its states in (sw_int16.cuh: T[i][j] = S[i][j] + i + j, so that gap extensions cost no
subtraction) gives the same last-column scan as the plain recurrences of smith_waterman_align (pemapper.c:1694-1748)
in units of 1/36: score, row, state and the tie flag, on random reads against random and related windows."""
import numpy as np

GO, GE, MATCH, MISM = 72, 1, 36, -12


def plain(read, win):
    mm, nn = len(read), len(win)
    NEG = -10**6
    S0 = np.full((nn + 1, mm + 1), NEG); S1 = S0.copy(); S2 = S0.copy()
    for j in range(1, mm + 1):
        S0[0][j] = S1[0][j] = S2[0][j] = -(GO + (j - 1) * GE)
    S0[:, 0] = 0; S1[:, 0] = 0; S2[:, 0] = -GO          # M[i][0] = 0 (2062-2081)
    best, bk, bi, tie = -(GO + (mm - 1) * GE), 0, 0, 0
    for i in range(1, nn + 1):
        for j in range(1, mm + 1):
            S2[i][j] = max(S0[i][j - 1] - GO, S2[i][j - 1] - GE)
            S1[i][j] = max(S0[i - 1][j] - GO, S1[i - 1][j] - GE)
            m = max(S0[i - 1][j - 1], S1[i - 1][j - 1], S2[i - 1][j - 1]) if j > 1 else 0
            S0[i][j] = m + (MATCH if read[j - 1] == win[i - 1] else MISM)
        for k, v in enumerate((S0[i][mm], S1[i][mm], S2[i][mm])):
            if v > best:
                best, bk, bi, tie = v, k, i, 0
            elif v == best:
                tie = 1
    return best, bk, bi, tie


def shifted(read, win):
    mm, nn = len(read), len(win)
    t0u = [-71] * (mm + 1); t1u = [-71] * (mm + 1); mu = [-81] * (mm + 1)   # row 0: T = -71, TM - 10 = -81
    best, bk, bi, tie = -71, 0, 0, 0                                          # S[0][0][mm] + mm
    for i in range(1, nn + 1):
        l0, l2, diag = i, i - 72, i - 11                                      # column 0 of row i
        for j in range(1, mm + 1):
            t2 = max(l0 - 71, l2)
            t1 = max(t0u[j] - 71, t1u[j])
            t0 = diag + (48 if read[j - 1] == win[i - 1] else 0)
            diag = mu[j]
            m = max(t0, t1, t2)
            t0u[j], t1u[j], mu[j] = t0, t1, m - 10
            l0, l2 = t0, t2
        for k, v in enumerate((t0u[mm] - i, t1u[mm] - i, l2 - i)):
            if v > best:
                best, bk, bi, tie = v, k, i, 0
            elif v == best:
                tie = 1
    return best - mm, bk, bi, tie


def test_shifted_frame_equals_plain_recurrences():
    rng = np.random.default_rng(3)
    for t in range(60):
        mm = int(rng.integers(16, 60))
        win = rng.integers(0, 4, mm + int(rng.integers(0, 22)))
        if t % 3 == 0:
            read = rng.integers(0, 4, mm)
        else:                                   # a noisy copy of part of the window, with an indel now and then
            o = int(rng.integers(0, len(win) - mm + 1))
            read = win[o:o + mm].copy()
            flip = rng.random(mm) < 0.08
            read[flip] = rng.integers(0, 4, int(flip.sum()))
            if t % 3 == 2:
                cut = int(rng.integers(4, mm - 4))
                read = np.concatenate([read[:cut], read[cut + 1:], rng.integers(0, 4, 1)])
        assert plain(read, win) == shifted(read, win), t
