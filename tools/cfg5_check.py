#!/usr/bin/env python
"""BASELINE.json configs[4] (high-repeat genome) on one B200: 16 contigs x 4 Mb, half of the sequence made of copies
of 200 2-kb units at 0-2 % divergence (many 16-mers with 2..99 positions, some >= 100), 150-bp reads of the cfg2
error model, paired and single-end.  Reports rate, stage times, candidates per read-mate, mapping types, fp64 share.
Parity on this genome family is pinned by the `repeat` fixture (tests/); this script is the throughput side.
Run on a B200:  python tools/cfg5_check.py [pairs] > profiles/cfg5_check.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pecaller_b200 as pb  # noqa: E402
from pecaller_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    dev = torch.device("cuda", 0)
    contigs = synth.repeat_genome(50, [4_000_000] * 16, unit_len=2000, n_units=200, frac=0.5, max_div=0.02)
    gt = torch.from_numpy(np.concatenate(contigs)).to(dev)
    d_r1, d_r2 = bench.torch_reads(gt, n, 51, dev)
    d_len = torch.full((n,), bench.READ_LEN, dtype=torch.int32, device=dev)
    m1 = torch.zeros(n, dtype=torch.int32, device=dev)
    m2 = torch.zeros(n, dtype=torch.int32, device=dev)
    ty = torch.zeros(n, dtype=torch.int32, device=dev)
    out = {"genome": "16 x 4 Mb, 50 % repeats of 200 x 2 kb units at 0-2 % divergence", "reads_per_run": n, "runs": []}
    for paired in (1, 0):
        params = pb.default_params(min_align=bench.MIN_ALIGN, pair_flag=paired, min_dist=bench.MIN_DIST, max_dist=bench.MAX_DIST)
        mapper = pb.PEMapper.from_genome(contigs, params, device=0)
        torch.cuda.synchronize()
        for it in range(3):
            mapper.reset_counts()
            mapper.reset_stats()
            mapper.map_device(n, d_r1.data_ptr(), d_len.data_ptr(), d_r2.data_ptr() if paired else 0,
                              d_len.data_ptr() if paired else 0, bench.STRIDE, bench.READ_LEN, m1.data_ptr(), m2.data_ptr(),
                              ty.data_ptr())
        st = mapper.stats()
        mates = n * (2 if paired else 1)
        out["runs"].append({"paired": bool(paired), "read_mates": mates, "reads_per_s": mates / (st["ms_total"] / 1e3),
                            "stage_ms": {k: st[k] for k in ("ms_seed", "ms_sw", "ms_select", "ms_tb_diag", "ms_tb_int",
                                                             "ms_tb_fp64", "ms_total")},
                            "candidates_per_mate": st["candidates"] / mates, "positions_per_mate": st["mer_positions"] / mates,
                            "replayed_fp64": st["replayed"], "fp64_tracebacks": st["exact_traced"],
                            "mapping_types": torch.bincount(ty.long(), minlength=9).tolist()})
        mapper.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
