#!/usr/bin/env python
"""BASELINE.json configs[3] as written: the Smith-Waterman kernel alone, reads of 100 / 150 / 250 bp (derived from the
window with 2 % edits) against 1000-bp random windows, GCUPS per read length on one B200 (pemap_sw_score_device ->
k_sw_i16<..., ROWCAP>), next to the in-envelope shape window = len + 21.  The reference cannot run 1000-bp windows (its
DP buffers are 300 x 300), so a sample of the results is checked against the oracle's restatement with enlarged buffers
(oracle/pemap_oracle.c: orc_sw_align_long); the same check over all edge shapes is tests/test_gpu_parity.py::test_cfg4.
    python tools/cfg4_sweep.py > profiles/cfg4_sweep_r02.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pecaller_b200 as pb  # noqa: E402


def make_pairs(gt, n, L, window, seed, dev):
    """windows at random genome offsets; the read is a stretch of its window with 2 % edits (substitutions, and 1-base
    insertions / deletions at 0.2 %)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    G = gt.shape[0]
    ws = torch.randint(0, G - window - 8, (n,), generator=g, device=dev)
    off = torch.randint(0, window - L - 4, (n,), generator=g, device=dev)
    step = torch.ones((n, L), dtype=torch.int64, device=dev)
    u = torch.rand((n, L), generator=g, device=dev)
    step += (u < 0.002)                      # deletion: skip a window base
    ins = u > 0.998                          # insertion: a random base, the window does not advance
    step[ins] = 0
    step[:, 0] = 0
    idx = (ws + off)[:, None] + torch.cumsum(step, dim=1)
    rows = gt[idx.clamp_(max=G - 1)]
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    rnd = acgt[torch.randint(0, 4, (n, L), generator=g, device=dev)]
    rows = torch.where(ins | (torch.rand((n, L), generator=g, device=dev) < 0.016), rnd, rows)
    stride = (L + 15) // 16 * 16
    buf = torch.zeros((n, stride), dtype=torch.uint8, device=dev)
    buf[:, :L] = rows
    return buf, stride, ws.to(torch.int32), torch.full((n,), window, dtype=torch.int32, device=dev)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dev = torch.device("cuda", 0)
    G = 16_000_000
    rng = np.random.Generator(np.random.PCG64(40))
    genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=G, dtype=np.uint8)]
    mapper = pb.PEMapper.from_genome([genome], pb.default_params(min_align=0.85, pair_flag=0), device=0)
    gt = torch.from_numpy(genome).to(dev)
    peak = json.load(open(os.path.join(ROOT, "profiles", "alu_peak.json")))["sw_s16x2_gcups_peak"]
    import oracle_lib as ol
    oracle = ol.Oracle([genome])
    out = []
    for L in (100, 150, 250):
        for window in (1000, L + 21):
            buf, stride, ws, wl = make_pairs(gt, n, L, window, 400 + L, dev)
            lens = torch.full((n,), L, dtype=torch.int32, device=dev)
            sc = torch.zeros(n, dtype=torch.int32, device=dev)
            mi = torch.zeros(n, dtype=torch.int32, device=dev)
            mk = torch.zeros(n, dtype=torch.int32, device=dev)
            fl = torch.zeros(n, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            ms = [mapper.sw_score_device(n, buf.data_ptr(), lens.data_ptr(), stride, L, ws.data_ptr(), wl.data_ptr(), window,
                                         sc.data_ptr(), mi.data_ptr(), mk.data_ptr(), fl.data_ptr()) for _ in range(5)]
            best = min(ms[2:])
            cells = float(n) * L * window
            # oracle check on a sample (the exact tie cases are flagged: there the double scan may pick another row)
            h_buf, h_ws, h_sc, h_mi, h_mk, h_fl = (t.cpu().numpy() for t in (buf, ws, sc, mi, mk, fl))
            bad = 0
            for i in range(0, n, max(1, n // 300)):
                s, (k, ii, _) = oracle.sw_align_long(int(h_ws[i]), window, h_buf[i, :L].tobytes())
                if round(s * 36) != int(h_sc[i]) or (not (h_fl[i] & 1) and (ii != int(h_mi[i]) or k != int(h_mk[i]))):
                    bad += 1
            out.append({"read_len": L, "window": window, "pairs": n, "cells": cells, "ms": best, "gcups": cells / best / 1e6,
                        "frac_of_s16x2_model_peak": cells / best / 1e6 / peak, "oracle_sample_mismatches": bad,
                        "mean_score": float(sc.float().mean().item()) / 36.0})
    oracle.close()
    mapper.close()
    print(json.dumps({"kernel": "k_sw_i16 (s16x2 DPX), window rows streamed through the wavefront", "peak_gcups_model": peak,
                      "sweep": out}, indent=1))


if __name__ == "__main__":
    main()
