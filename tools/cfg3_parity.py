#!/usr/bin/env python
"""BASELINE.md section 3.4 on the GPU box: a prefix of cfg3 (3.1 Gb, 24 contigs, 150 bp pairs of the cfg2 error model)
mapped by the CUDA path and by the UNMODIFIED reference (oracle/_ref/libpemapper_ref.so = pemapper.c's own
map_everything on all host threads), compared bit for bit: .mfile values (m1/m2) and mapping type of every read,
sha256 of the pileup records, insertion multiset, and the nine type counts of summary.txt.  Also times both.
Needs ~160 GB of host RAM (the reference's BASE_NODE array is 32 B per genome base).

    python tools/cfg3_parity.py [pairs] > profiles/cfg3_parity_r02.json          (pairs >= 2,000,000 for the record)

TEST INFRASTRUCTURE: the reference library is the checker here, never part of the product path."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import oracle_lib as ol  # noqa: E402
import pecaller_b200 as pb  # noqa: E402


def run(n_pairs, config="cfg3", threads=None, log=sys.stderr, paired=True, also_single=0):
    """paired / single-end run of n_pairs; also_single > 0: after a paired run, that many single-end reads (mate 1 of the
    same pairs) through the same index, mapper and reference instance (one process holds one reference genome)."""
    threads = threads or (os.cpu_count() or 1)
    dev = torch.device("cuda", 0)
    cfg = bench.CONFIGS[config]
    t0 = time.time()
    contigs, gt = bench.config_genome(config, dev)
    d_r1, d_r2 = bench.torch_reads(gt, max(n_pairs, also_single), cfg["read_seed"], dev)
    r1_all = d_r1[:, :bench.READ_LEN].cpu().numpy()
    r2_all = d_r2[:, :bench.READ_LEN].cpu().numpy()
    del gt, d_r1, d_r2
    torch.cuda.empty_cache()
    params = pb.default_params(min_align=bench.MIN_ALIGN, pair_flag=int(paired), min_dist=bench.MIN_DIST, max_dist=bench.MAX_DIST)
    pos_index, mers, cstarts = bench.host_index(contigs, params, True)
    print("data + host index: %.1f s" % (time.time() - t0), file=log, flush=True)
    t0 = time.time()
    mapper = pb.PEMapper.from_genome(contigs, params)
    t_init = time.time() - t0
    t0 = time.time()
    cat = np.concatenate(contigs) if len(contigs) > 1 else contigs[0]
    ref = ol.ReferenceLib(None, min_align=bench.MIN_ALIGN, paired=paired, min_dist=bench.MIN_DIST, max_dist=bench.MAX_DIST,
                          arrays=(cat, cstarts, pos_index, mers, len(contigs)))
    t_ref_init = time.time() - t0

    def one(n, is_paired):
        mates = 2 if is_paired else 1
        r1 = r1_all[:n]
        r2 = r2_all[:n] if is_paired else None
        # ---- CUDA path through the C-ABI (host rows in, records through the bounded finish)
        mapper.set_params(pair_flag=int(is_paired))
        mapper.reset_counts()
        mapper.reset_stats()
        t0 = time.time()
        gm1, gm2, gty = mapper.map_batch(r1, r2)
        t_gpu = time.time() - t0
        st = mapper.stats()
        h = hashlib.sha256()
        acc = {"n": 0}

        def consume(rec):
            h.update(rec.tobytes())
            acc["n"] += rec.shape[0]
        mapper.finish_stream(consume)
        gins = sorted(mapper.insertions())
        gpu = {"records": acc["n"], "records_sha256": h.hexdigest(), "insertions": len(gins),
               "insertions_sha256": hashlib.sha256(repr(gins).encode()).hexdigest(),
               "type_counts": np.bincount(gty, minlength=9).tolist(), "map_batch_s": round(t_gpu, 3),
               "reads_per_s_host_rows": mates * n / t_gpu, "reads_per_s_kernels": mates * n / (st["ms_total"] / 1e3),
               "replayed_fp64_read_mates": int(st["replayed"]), "fp64_tracebacks": int(st["exact_traced"]),
               "init_s": round(t_init, 1)}
        print("CUDA path: %.2f s for %d %s" % (t_gpu, n, "pairs" if is_paired else "reads"), file=log, flush=True)
        # ---- the unmodified reference on all host threads
        ref.set_params(bench.MIN_ALIGN, is_paired, bench.MIN_DIST, bench.MAX_DIST)
        ref.reset()
        t0 = time.time()
        rm1, rm2, rty = ref.map(r1, r2, nthreads=threads)
        t_ref = time.time() - t0
        print("reference: %.1f s on %d threads" % (t_ref, threads), file=log, flush=True)
        rrec = ref.records()
        rins = ref.insertions()
        cpu = {"records": int(rrec.shape[0]), "records_sha256": hashlib.sha256(rrec.tobytes()).hexdigest(),
               "insertions": len(rins), "insertions_sha256": hashlib.sha256(repr(rins).encode()).hexdigest(),
               "type_counts": np.bincount(rty, minlength=9).tolist(), "map_s": round(t_ref, 2), "reads_per_s": mates * n / t_ref,
               "threads": threads, "threads_live": ref.live_threads(), "init_s": round(t_ref_init, 1)}
        same = {"m1": bool(np.array_equal(gm1, rm1)), "m2": bool(np.array_equal(gm2, rm2)) if is_paired else True,
                "mapping_type": bool(np.array_equal(gty, rty)),
                "pileup_records": gpu["records"] == cpu["records"] and gpu["records_sha256"] == cpu["records_sha256"],
                "insertions": gins == rins, "type_counts": gpu["type_counts"] == cpu["type_counts"]}
        return {"config": bench.workload_text(config, contigs, n).replace(" per GPU per step", " (prefix, max_reads)")
                .replace("paired-end", "paired-end" if is_paired else "single-end"),
                "pairs": n, "paired": bool(is_paired), "cuda": gpu, "reference": cpu, "identical": same,
                "all_identical": all(same.values()),
                "speedup_host_rows_vs_reference": gpu["reads_per_s_host_rows"] / cpu["reads_per_s"]}

    out = one(n_pairs, paired)
    if also_single:
        out["single_end"] = one(also_single, False)
        out["all_identical"] = out["all_identical"] and out["single_end"]["all_identical"]
    mapper.close()
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    out = run(n, os.environ.get("PEMAP_PARITY_CONFIG", "cfg3"), paired=os.environ.get("PEMAP_PARITY_SINGLE", "0") != "1",
              also_single=int(os.environ.get("PEMAP_PARITY_ALSO_SINGLE", "0")))
    print(json.dumps(out, indent=1))
    sys.exit(0 if out["all_identical"] else 1)
