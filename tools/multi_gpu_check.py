#!/usr/bin/env python
"""N GPUs == 1 GPU, byte for byte, at the real sizes (BASELINE configs[2] and configs[4]).

Launched like bench.py (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py cfg3 2000000 > profiles/multi_gpu_cfg3_r02.json
Every rank builds the same genome and index, maps its round-robin share of the same P pairs (pecaller_b200.sharding),
the pileup counters are summed slice-wise over NVLink peer memory (pemap_reduce_scatter_ipc), every rank compacts its
own slice and hashes its records.  Rank 0 then maps ALL the pairs alone on its GPU and the two results are compared:
m1 / m2 / mapping type of every pair, the sha256 of the concatenated pileup records, the insertion multiset."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pecaller_b200 as pb  # noqa: E402
from pecaller_b200 import sharding  # noqa: E402


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    paired = os.environ.get("PEMAP_CHECK_SINGLE", "0") != "1"
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    cfg = bench.CONFIGS[config]
    contigs, gt = bench.config_genome(config, dev)
    d_r1, d_r2 = bench.torch_reads(gt, n, cfg["read_seed"], dev)     # the same P pairs on every rank
    r1 = d_r1[:, :bench.READ_LEN].cpu().numpy()
    r2 = d_r2[:, :bench.READ_LEN].cpu().numpy() if paired else None
    del gt, d_r1, d_r2
    torch.cuda.empty_cache()
    params = pb.default_params(min_align=bench.MIN_ALIGN, pair_flag=int(paired), min_dist=bench.MIN_DIST, max_dist=bench.MAX_DIST)
    mapper = pb.PEMapper.from_genome(contigs, params, device=local)
    batch = 20_000                                                     # the reference's reads_per_thread
    ranges = sharding.shard_batches(n, rank, world, batch)
    t0 = time.time()
    parts = [mapper.map_batch(r1[a:b], r2[a:b] if paired else None) for a, b in ranges]
    m1, m2, ty = (np.concatenate([p[k] for p in parts]) for k in range(3))
    red = sharding.SliceReducer(mapper)
    lo, hi = red.reduce_scatter()
    recs = []
    mapper.finish_stream(lambda r: recs.append(r.copy()), site_range=(lo, hi))
    mine = np.concatenate(recs) if recs else np.zeros(0, dtype=pb.RECORD_DTYPE)
    t_multi = time.time() - t0
    ins = mapper.insertions()
    res = sharding.gather_results(n, ranges, m1, m2, ty, dst=0)
    slice_digest = hashlib.sha256(mine.tobytes()).hexdigest()
    all_dig = [None] * world
    all_n = [None] * world
    all_ins = [None] * world
    dist.all_gather_object(all_dig, slice_digest)
    dist.all_gather_object(all_n, int(mine.shape[0]))
    dist.all_gather_object(all_ins, ins)
    dist.barrier()
    if rank == 0:
        mapper.reset_counts()
        t0 = time.time()
        s1, s2, sty = mapper.map_batch(r1, r2 if paired else None)
        # the single-GPU records, cut at the same slice boundaries so that the digests compare slice by slice
        tile = 2048
        tiles = (int(sum(c.shape[0] for c in contigs)) + tile - 1) // tile
        per = (tiles + world - 1) // world
        G = int(sum(c.shape[0] for c in contigs))
        single_dig, single_n = [], []
        for r in range(world):
            a, b = min(G, r * per * tile), min(G, (r + 1) * per * tile)
            part = []
            mapper.finish_stream(lambda x: part.append(x.copy()), site_range=(a, b))
            arr = np.concatenate(part) if part else np.zeros(0, dtype=pb.RECORD_DTYPE)
            single_dig.append(hashlib.sha256(arr.tobytes()).hexdigest())
            single_n.append(int(arr.shape[0]))
        t_single = time.time() - t0
        sins = mapper.insertions()
        same = {"m1": bool(np.array_equal(res[0], s1)), "m2": bool(np.array_equal(res[1], s2)),
                "mapping_type": bool(np.array_equal(res[2], sty)), "pileup_slices": all_dig == single_dig and all_n == single_n,
                "insertions": sorted(sum(all_ins, [])) == sorted(sins)}
        out = {"config": bench.workload_text(config, contigs, n).replace(" per GPU per step", " in total"), "gpus": world,
               "paired": paired, "pairs": n, "records": int(sum(all_n)), "insertions": len(sins),
               "records_sha256_per_slice": all_dig, "identical_to_one_gpu": same, "all_identical": all(same.values()),
               "seconds_n_gpus_map_sum_compact": round(t_multi, 2), "seconds_one_gpu": round(t_single, 2),
               "type_counts": np.bincount(sty, minlength=9).tolist()}
        print(json.dumps(out, indent=1), flush=True)
    dist.barrier()
    mapper.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
