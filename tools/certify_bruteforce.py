#!/usr/bin/env python
"""Brute-force check of the ungapped-diagonal certificate (k_diag_certify, sw_int16.cuh) on the CPU.

Small random (window, read) cases over a 2-4 letter alphabet, so that coincidences and repeats are frequent: whenever
the certificate fires, the full three-state integer DP (units of 1/36, as k_sw_i16 computes it) must give the same
score / maxi / maxk, a unique last-column maximum and a traceback that is the pure state-0 diagonal decided by
strict inequalities.  Usage: python tools/certify_bruteforce.py [cases] [seed]
"""
import random
import sys


def dp(win, read):
    nn, mm = len(win), len(read)
    NEG = -10 ** 9
    S0 = [[NEG] * (mm + 1) for _ in range(nn + 1)]
    S1 = [[NEG] * (mm + 1) for _ in range(nn + 1)]
    S2 = [[NEG] * (mm + 1) for _ in range(nn + 1)]
    for i in range(nn + 1):
        S0[i][0] = 0
        S1[i][0] = 0
        S2[i][0] = -72
    for j in range(1, mm + 1):
        b = -(72 + j - 1)
        S0[0][j] = S1[0][j] = S2[0][j] = b
    for i in range(1, nn + 1):
        for j in range(1, mm + 1):
            S2[i][j] = max(S0[i][j - 1] - 72, S2[i][j - 1] - 1)
            S1[i][j] = max(S0[i - 1][j] - 72, S1[i - 1][j] - 1)
            m = max(S0[i - 1][j - 1], S1[i - 1][j - 1], S2[i - 1][j - 1])
            S0[i][j] = m + (36 if win[i - 1] == read[j - 1] else -12)
    # last-column scan (1717-1742): i ascending, states 0,1,2, strict >, starting from S0[0][mm]
    best, bk, bi, tie = S0[0][mm], 0, 0, False
    for i in range(1, nn + 1):
        for k, S in enumerate((S0, S1, S2)):
            v = S[i][mm]
            if v > best:
                best, bk, bi, tie = v, k, i, False
            elif v == best:
                tie = True
    # pure-diagonal certificate of k_sw_i16: S0 - max(S1, S2) >= 1 on every cell of the diagonal back from (bi, mm)
    pure = bk == 0 and bi > 0
    if pure:
        i, j = bi, mm
        while i >= 1 and j >= 1:
            if S0[i][j] - max(S1[i][j], S2[i][j]) < 1:
                pure = False
                break
            i -= 1
            j -= 1
    return best, bi, bk, tie, pure


def certify(win, read, max_m):
    nn, mm = len(win), len(read)
    K = nn - mm
    if K < 0 or mm < 8:
        return None
    good = []
    pre, suf = {}, {}
    for o in range(K + 1):
        mis = [j for j in range(mm) if win[o + j] != read[j]]
        p = 0
        while p < mm and win[o + p] == read[p]:
            p += 1
        s = 0
        while s < mm and win[o + mm - 1 - s] == read[mm - 1 - s]:
            s += 1
        pre[o], suf[o] = p, s
        if len(mis) <= max_m:
            good.append((o, len(mis), (mis[0] + 1) if mis else mm + 1))
    if len(good) != 1:
        return None
    o, m, p1 = good[0]
    if m == 2:
        A = max([pre[x] for x in pre if x != o] or [0])
        B = max([suf[x] for x in suf if x != o] or [0])
        if not (A <= p1 and B <= mm - p1 and A + B <= mm - 1):
            return None
    return 36 * mm - 48 * m, o + mm, 0


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    fired = {1: 0, 2: 0}
    for max_m in (1, 2):
        for _ in range(cases):
            alpha = "ACGT"[: rng.choice((2, 2, 3, 4))]
            mm = rng.randint(8, 20)
            K = rng.randint(0, 8)
            o = rng.randint(0, K)
            read = [rng.choice(alpha) for _ in range(mm)]
            win = [rng.choice(alpha) for _ in range(mm + K)]
            mode = rng.random()
            if mode < 0.8:      # plant the read on diagonal o with a few substitutions
                for j in range(mm):
                    win[o + j] = read[j]
                for _ in range(rng.choice((0, 1, 1, 2, 2, 2, 3))):
                    win[o + rng.randrange(mm)] = rng.choice(alpha)
            if mode < 0.3:      # and a tandem repeat / low-complexity background
                unit = [rng.choice(alpha) for _ in range(rng.randint(1, 4))]
                for x in range(len(win)):
                    if rng.random() < 0.7:
                        win[x] = unit[x % len(unit)]
                if rng.random() < 0.7:
                    for j in range(mm):
                        if rng.random() < 0.8:
                            read[j] = win[o + j]
            c = certify(win, read, max_m)
            if c is None:
                continue
            fired[max_m] += 1
            best, bi, bk, tie, pure = dp(win, read)
            if (best, bi, bk) != c or tie or not pure:
                print("MISMATCH max_m=%d" % max_m, "".join(win), "".join(read), c, (best, bi, bk, tie, pure))
                sys.exit(1)
    print("ok: certificate fired %d (m<=1 rule) + %d (m<=2 rule) times in %d cases each, all equal to the DP" %
          (fired[1], fired[2], cases))


if __name__ == "__main__":
    main()
