#!/usr/bin/env python
"""BASELINE.json configs[2] at genome scale on ONE B200: 24 contigs (sizes proportional to human chr1-22, X, Y)
totalling 3.1e9 bp, index built on the device, N pairs of 150 bp reads of the cfg2 error model drawn from it.
Reports the device-resident mapping rate, stage times, mapping-type counts, HBM footprint and the pileup size.
(The 30x = 310 M pairs of the config text are 99 GB of reads; the rate per read is what this measures.)
Run on a B200:  python tools/cfg3_check.py [pairs] > profiles/cfg3_check.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pecaller_b200 as pb  # noqa: E402

CHR_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    total = float(os.environ.get("PEMAP_CFG3_BASES", 3.1e9))
    dev = torch.device("cuda", 0)
    lens = [int(total * m / sum(CHR_MB)) for m in CHR_MB]
    G = sum(lens)
    t0 = time.time()
    g = torch.Generator(device=dev)
    g.manual_seed(30)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    gt = torch.empty(G, dtype=torch.uint8, device=dev)
    step = 1 << 28
    for lo in range(0, G, step):
        m = min(step, G - lo)
        gt[lo:lo + m] = acgt[torch.randint(0, 4, (m,), generator=g, device=dev)]
    host = gt.cpu().numpy()
    t_gen = time.time() - t0
    contigs, at = [], 0
    for L in lens:
        contigs.append(host[at:at + L])
        at += L
    params = pb.default_params(min_align=bench.MIN_ALIGN, pair_flag=1, min_dist=bench.MIN_DIST, max_dist=bench.MAX_DIST)
    t0 = time.time()
    mapper = pb.PEMapper.from_genome(contigs, params, device=0)
    t_index = time.time() - t0
    free_b, total_b = torch.cuda.mem_get_info()
    d_r1, d_r2 = bench.torch_reads(gt, n, 31, dev)
    d_len = torch.full((n,), bench.READ_LEN, dtype=torch.int32, device=dev)
    m1 = torch.zeros(n, dtype=torch.int32, device=dev)
    m2 = torch.zeros(n, dtype=torch.int32, device=dev)
    ty = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for it in range(4):  # 3 warm-up passes, the 4th is measured
        mapper.reset_counts()
        mapper.reset_stats()
        mapper.map_device(n, d_r1.data_ptr(), d_len.data_ptr(), d_r2.data_ptr(), d_len.data_ptr(), bench.STRIDE,
                          bench.READ_LEN, m1.data_ptr(), m2.data_ptr(), ty.data_ptr())
    st = mapper.stats()
    t0 = time.time()
    acc = {"records": 0, "counted": 0}

    def consume(r):  # what a writer would do: look at every record once
        acc["records"] += r.shape[0]
        acc["counted"] += int(r["c"].sum(dtype=np.int64))
    n_rec = mapper.finish_stream(consume)
    t_finish = time.time() - t0
    t0 = time.time()
    n_rec2 = mapper.finish_stream(None)   # bounded compaction + pinned D2H alone
    t_finish_raw = time.time() - t0
    ins = mapper.insertions()
    free_b2, _ = torch.cuda.mem_get_info()
    out = {"genome_bases": G, "contigs": len(lens), "pairs": n, "datagen_s": round(t_gen, 1), "index_build_s": round(t_index, 1),
           "hbm_used_gb_after_init": round((total_b - free_b) / 1e9, 1), "reads_per_s": 2 * n / (st["ms_total"] / 1e3),
           "stage_ms": {k: st[k] for k in ("ms_seed", "ms_sw", "ms_select", "ms_traceback", "ms_tb_diag", "ms_tb_int",
                                            "ms_tb_fp64", "ms_total")},
           "mapped_read_mates": int((m1 != 0).sum().item() + (m2 != 0).sum().item()),
           "mapping_types": torch.bincount(ty.long(), minlength=9).tolist(),
           "candidates": st["candidates"], "lookups": st["lookups"], "mer_positions": st["mer_positions"],
           "seed_glookups_s": st["lookups"] / (st["ms_seed"] / 1e3) / 1e9,
           "pileup_records": int(n_rec), "counted": acc["counted"], "insertions": len(ins),
           "finish_stream_s": round(t_finish, 2), "finish_stream_d2h_only_s": round(t_finish_raw, 3),
           "finish_d2h_gbs": round(16 * n_rec2 / max(t_finish_raw, 1e-9) / 1e9, 1),
           "hbm_used_gb_after_finish": round((total_b - free_b2) / 1e9, 1)}
    print(json.dumps(out, indent=1))
    mapper.close()


if __name__ == "__main__":
    main()
