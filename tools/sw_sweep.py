#!/usr/bin/env python
"""SW kernel sweep (BASELINE.json configs[3], in-envelope shape): single-end reads of 100 / 150 / 250 bp with 2 %
substitutions against their len+21 windows (the reference's own window, pemapper.c:1047-1081; its 300x300 DP buffers
rule out the 1000-bp windows of the config text).  Reports the integer scoring kernel's GCUPS per read length from the
library's own CUDA-event stage times and cell counters.  The ungapped-diagonal certificate (k_diag_certify) is switched
off (PEMAP_CERTIFY=0) so that every candidate goes through the DP kernel: this is the kernel microbenchmark; the second
figure per read length is the stage with the certificate on (effective rate over all candidates' cells).
Run on a B200:  python tools/sw_sweep.py > profiles/sw_sweep.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pecaller_b200 as pb  # noqa: E402


def main():
    os.environ["PEMAP_CERTIFY"] = "0"
    kernel = run(certify=False)
    os.environ["PEMAP_CERTIFY"] = "1"
    eff = run(certify=True)
    for k, e in zip(kernel["sweep"], eff["sweep"]):
        k["with_certificate"] = {"ms_sw": e["ms_sw"], "effective_gcups_all_cells": e["sw_gcups_all_cells"],
                                 "certified_cell_frac": e["certified_cell_frac"], "reads_per_s": e["reads_per_s"]}
    print(json.dumps(kernel, indent=1))


def run(certify):
    dev = torch.device("cuda", 0)
    G = 16_000_000
    rng = np.random.Generator(np.random.PCG64(40))
    genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=G, dtype=np.uint8)]
    mapper = pb.PEMapper.from_genome([genome], pb.default_params(min_align=0.85, pair_flag=0), device=0)
    gt = torch.from_numpy(genome).to(dev)
    comp = torch.full((256,), ord("N"), dtype=torch.uint8, device=dev)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    peak = json.load(open(os.path.join(ROOT, "profiles", "alu_peak.json")))["sw_s16x2_gcups_peak"]
    out = []
    n = 2_000_000
    for L in (100, 150, 250):
        g = torch.Generator(device=dev)
        g.manual_seed(40 + L)
        stride = (L + 15) // 16 * 16
        start = torch.randint(0, G - L - 1, (n,), generator=g, device=dev)
        rows = gt[start[:, None] + torch.arange(L, device=dev)[None, :]]
        sub = torch.rand((n, L), generator=g, device=dev) < 0.02
        rows = torch.where(sub, acgt[torch.randint(0, 4, (n, L), generator=g, device=dev)], rows)
        rev = torch.rand((n,), generator=g, device=dev) < 0.5
        rows = torch.where(rev[:, None], comp[rows.flip(1).long()], rows)
        buf = torch.zeros((n, stride), dtype=torch.uint8, device=dev)
        buf[:, :L] = rows
        lens = torch.full((n,), L, dtype=torch.int32, device=dev)
        m1 = torch.zeros(n, dtype=torch.int32, device=dev)
        m2 = torch.zeros(n, dtype=torch.int32, device=dev)
        ty = torch.zeros(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()  # the library runs on its own non-blocking stream
        for it in range(4):  # 3 warm-up passes, the 4th is measured
            mapper.reset_counts()
            mapper.reset_stats()
            mapper.map_device(n, buf.data_ptr(), lens.data_ptr(), 0, 0, stride, L, m1.data_ptr(), m2.data_ptr(), ty.data_ptr())
        st = mapper.stats()
        dp_cells = st["sw_cells"] - st["sw_cells_certified"]
        gc = dp_cells / (st["ms_sw"] / 1e3) / 1e9
        out.append({"read_len": L, "window": L + 21, "reads": n, "candidates": st["candidates"], "sw_cells": st["sw_cells"],
                    "ms_sw": st["ms_sw"], "sw_gcups": gc, "frac_of_s16x2_model_peak": gc / peak,
                    "sw_gcups_all_cells": st["sw_cells"] / (st["ms_sw"] / 1e3) / 1e9,
                    "certified_cell_frac": st["sw_cells_certified"] / max(st["sw_cells"], 1),
                    "mapped": int((m1 != 0).sum().item()), "ms_seed": st["ms_seed"], "ms_traceback": st["ms_traceback"],
                    "reads_per_s": n / (st["ms_total"] / 1e3)})
    mapper.close()
    return {"kernel": "k_sw_i16 (s16x2 DPX), certificate off", "peak_gcups_model": peak, "sweep": out}


if __name__ == "__main__":
    main()
