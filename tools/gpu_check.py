#!/usr/bin/env python
"""Diagnostic parity run on a GPU box: CUDA path (through the C-ABI) vs the oracle, stage by stage, with a
mismatch dump into gpurun_out/.  tests/test_gpu_parity.py is the formal version; this prints more.

Usage: python tools/gpu_check.py [--max-reads N] fixture [fixture...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures_def  # noqa: E402
import oracle_lib as ol  # noqa: E402
import pecaller_b200 as pb  # noqa: E402


def compare_run(mapper, oracle, run, max_reads, out):
    r1 = run.reads1[:max_reads]
    r2 = run.reads2[:max_reads] if run.paired else None
    kw = dict(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist)
    oracle.reset()
    oracle.set_params(**kw)
    mapper.reset_counts()
    mapper.reset_stats()
    mapper.set_params(**kw)
    mapper.keep(pb.KEEP_DETAIL | pb.KEEP_CANDIDATES)
    t = time.time()
    om1, om2, oty, odet = oracle.map_batch(r1, r2, nthreads=os.cpu_count(), detail=True)
    t_or = time.time() - t
    t = time.time()
    gm1, gm2, gty = mapper.map_batch(r1, r2)
    t_gpu = time.time() - t
    gdet = mapper.detail(r1.shape[0])
    res = {"run": run.name, "n": int(r1.shape[0]), "oracle_s": round(t_or, 3), "gpu_s": round(t_gpu, 3)}
    res["m1_equal"] = bool(np.array_equal(om1, gm1))
    res["m2_equal"] = bool(np.array_equal(om2, gm2))
    res["type_equal"] = bool(np.array_equal(oty, gty))
    for f in ("hits1", "hits2", "best1", "best2", "orient1", "orient2"):
        res[f + "_equal"] = bool(np.array_equal(odet[f], gdet[f]))
    for f in ("score1", "score2"):
        res[f + "_bits_equal"] = bool(np.array_equal(odet[f].view(np.uint64), gdet[f].view(np.uint64)))
    bad = np.nonzero((om1 != gm1) | (om2 != gm2) | (oty != gty) | (odet["hits1"] != gdet["hits1"]) |
                     (odet["hits2"] != gdet["hits2"]))[0]
    res["n_bad_reads"] = int(bad.shape[0])
    dump = []
    for i in bad[:20]:
        d = {"i": int(i), "oracle": [int(om1[i]), int(om2[i]), int(oty[i])], "gpu": [int(gm1[i]), int(gm2[i]), int(gty[i])],
             "odet": [x.item() for x in odet[i]], "gdet": [x.item() for x in gdet[i]], "read1": r1[i].tobytes().decode()}
        os_, oo = oracle.initial_map(r1[i].tobytes())
        gs, go = mapper.candidates(int(i), 0)
        d["ocand1"] = [os_.tolist()[:12], oo.tolist()[:12]]
        d["gcand1"] = [gs.tolist()[:12], go.tolist()[:12]]
        dump.append(d)
    # candidate lists of a sample of reads
    n_c = 0
    cand_bad = 0
    for i in range(0, r1.shape[0], max(1, r1.shape[0] // 300)):
        os_, oo = oracle.initial_map(r1[i].tobytes())
        gs, go = mapper.candidates(i, 0)
        n_c += 1
        if not (np.array_equal(os_, gs) and np.array_equal(oo, go)):
            cand_bad += 1
    res["cand_checked"] = n_c
    res["cand_bad"] = cand_bad
    orec = oracle.records()
    grec, gins = mapper.finish()
    res["records_equal"] = bool(orec.shape == grec.shape and orec.tobytes() == grec.tobytes())
    res["n_records"] = [int(orec.shape[0]), int(grec.shape[0])]
    oins = oracle.insertions()
    res["insertions_equal"] = bool(oins == sorted(gins))
    res["n_insertions"] = [len(oins), len(gins)]
    if not res["records_equal"] and orec.shape == grec.shape:
        w = np.nonzero((orec["pos"] != grec["pos"]) | (orec["c"] != grec["c"]).any(axis=1))[0]
        res["first_bad_records"] = [[orec[j].tolist(), grec[j].tolist()] for j in w[:5]]
    res["stats"] = mapper.stats()
    res["oracle_cells"] = int(oracle.cells())
    # integer fast path (default mode)
    mapper.keep(0)
    mapper.reset_counts()
    mapper.reset_stats()
    fm1, fm2, fty = mapper.map_batch(r1, r2)
    frec, fins = mapper.finish()
    res["fast_m1_equal"] = bool(np.array_equal(om1, fm1))
    res["fast_m2_equal"] = bool(np.array_equal(om2, fm2))
    res["fast_type_equal"] = bool(np.array_equal(oty, fty))
    res["fast_records_equal"] = bool(orec.shape == frec.shape and orec.tobytes() == frec.tobytes())
    res["fast_insertions_equal"] = bool(oins == sorted(fins))
    fst = mapper.stats()
    res["fast_stats"] = {k: fst[k] for k in ("reads", "replayed", "ms_seed", "ms_sw", "ms_select", "ms_traceback", "ms_total")}
    fbad = np.nonzero((om1 != fm1) | (om2 != fm2) | (oty != fty))[0]
    res["fast_bad_reads"] = [[int(i), int(om1[i]), int(om2[i]), int(oty[i]), int(fm1[i]), int(fm2[i]), int(fty[i])] for i in fbad[:10]]
    out.append({"summary": res, "bad": dump})
    print(json.dumps(res), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-reads", type=int, default=1 << 30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_check.json"))
    ap.add_argument("fixtures", nargs="+")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    out = []
    ok = True
    for name in a.fixtures:
        fx = fixtures_def.FIXTURES[name]()
        t = time.time()
        oracle = ol.Oracle(fx.genome)
        t1 = time.time()
        mapper = pb.PEMapper.from_genome(fx.genome)
        t2 = time.time()
        print("fixture %s: oracle index %.1fs, device init+index %.1fs" % (name, t1 - t, t2 - t1), flush=True)
        # index parity: mers byte for byte, pos_index at every distinct k-mer boundary region (sampled)
        gm = mapper.read_mers()
        omers = oracle.mers()
        idx_ok = bool(np.array_equal(gm, omers))
        rng = np.random.default_rng(0)
        ws = rng.integers(0, 1 << 32, size=2000, dtype=np.uint64)
        pi_ok = all(int(mapper.read_pos_index(int(w), 2)[0]) == oracle.pos_index(int(w)) for w in ws[:500])
        tail = mapper.read_pos_index((1 << 32) - 1, 2)
        pi_ok = pi_ok and int(tail[1]) == omers.shape[0] and int(tail[0]) == oracle.pos_index((1 << 32) - 1)
        print(json.dumps({"fixture": name, "mers_equal": idx_ok, "pos_index_sample_equal": pi_ok, "n_mers": int(gm.shape[0])}),
              flush=True)
        ok = ok and idx_ok and pi_ok
        for run in fx.runs:
            r = compare_run(mapper, oracle, run, a.max_reads, out)
            ok = ok and all(v for k, v in r.items() if k.endswith("_equal")) and r["cand_bad"] == 0
        mapper.close()
        oracle.close()
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    print("ALL OK" if ok else "MISMATCHES", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
