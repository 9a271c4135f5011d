#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference programs (oracle/_ref, built from
/root/reference/src by oracle/Makefile) on the seeded fixtures of tests/fixtures_def.py.

Runs only in the build container (needs /root/reference, ~20 GB RAM, minutes per fixture).
Usage: python tools/make_golden.py [--work /tmp/golden] [--threads 8] fixture [fixture...]

For every run of every fixture it stores under tests/golden/<fixture>/:
  <run>.mfile1.gz / .mfile2.gz   raw uint32 per read (reference .mfile), gzip
  <run>.summary.txt              reference .summary.txt verbatim
  <run>.pileup.json              {n_records, sha256 of the inflated .pileup.gz, per-column sums}
  <run>.pileup.bin.gz            the inflated pileup itself (only when small)
  <run>.indel.norm.txt.gz        .indel.txt.gz with the insertion strings of each line sorted
plus <fixture>/index.json: sha256 of .mdx, .sdx text, sha256 of the inflated .idx (16 GiB stream).
"""
import argparse
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pecaller_b200 import synth  # noqa: E402
import fixtures_def  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def sha_stream(opener, path, chunk=1 << 26):
    h = hashlib.sha256()
    n = 0
    with opener(path, "rb") as f:
        while True:
            b = f.read(chunk)
            if not b:
                break
            h.update(b)
            n += len(b)
    return h.hexdigest(), n


def build_index(work, fx):
    fa = os.path.join(work, "g.fa")
    if not os.path.exists(os.path.join(work, "g.sdx")):
        synth.write_fasta(fa, fx.genome, fx.names)
        t = time.time()
        stdin = "S\n%d\ng.fa\ng\n%s\n" % (len(fx.genome) + 1, "y" if getattr(fx, "bisulfite", False) else "n")
        subprocess.run([os.path.join(REF, "index_genome_whole")], input=stdin.encode(), cwd=work,
                       stdout=subprocess.DEVNULL, check=True)
        print("  index_genome_whole: %.0f s" % (time.time() - t), flush=True)


def run_mapper(work, fx, run, threads):
    out = run.name
    f1 = os.path.join(work, run.name + "_1.fq")
    synth.write_fastq(f1, run.reads1)
    n = run.reads1.shape[0]
    bis = "y" if run.bisulfite else "n"
    if run.paired:
        f2 = os.path.join(work, run.name + "_2.fq")
        synth.write_fastq(f2, run.reads2)
        cmd = ["pemapper", out, "g.sdx", "p", run.name + "_1.fq", run.name + "_2.fq", str(run.max_dist),
               str(run.min_dist), bis, repr(run.min_align), str(threads), str(n + 8)]
    else:
        cmd = ["pemapper", out, "g.sdx", "s", run.name + "_1.fq", bis, repr(run.min_align), str(threads), str(n + 8)]
    t = time.time()
    cmd[0] = os.path.join(REF, "pemapper")
    subprocess.run(cmd, cwd=work, stdout=subprocess.DEVNULL, check=True)
    print("  pemapper %s: %.0f s  (%s)" % (run.name, time.time() - t, " ".join(cmd[1:])), flush=True)
    return cmd


def normalise_indel(path):
    lines = []
    with gzip.open(path, "rt") as f:
        for ln in f.read().split("\n"):
            parts = ln.split("\t")
            if len(parts) > 7 and parts[0] != "Fragment":
                parts = parts[:7] + sorted(parts[7:])
            lines.append("\t".join(parts))
    return "\n".join(lines)


def collect(work, gold, fx, run, cmd):
    n = run.reads1.shape[0]
    for k, suffix in ((1, "_1.fq.mfile"), (2, "_2.fq.mfile")):
        p = os.path.join(work, run.name + suffix)
        if os.path.exists(p):
            raw = np.fromfile(p, dtype=np.uint32)
            assert raw.shape[0] == n, (raw.shape, n)
            with gzip.open(os.path.join(gold, "%s.mfile%d.gz" % (run.name, k)), "wb", compresslevel=9) as f:
                f.write(raw.tobytes())
    shutil.copy(os.path.join(work, run.name + ".summary.txt"), os.path.join(gold, run.name + ".summary.txt"))
    with gzip.open(os.path.join(work, run.name + ".pileup.gz"), "rb") as f:
        raw = f.read()
    rec = np.frombuffer(raw, dtype=np.dtype([("pos", "<u4"), ("c", "<u2", (6,))]))
    meta = {"n_records": int(rec.shape[0]), "sha256": hashlib.sha256(raw).hexdigest(),
            "column_sums": [int(x) for x in rec["c"].astype(np.int64).sum(axis=0)],
            "pos_sum": int(rec["pos"].astype(np.int64).sum()), "cmd": cmd[1:]}
    with open(os.path.join(gold, run.name + ".pileup.json"), "w") as f:
        json.dump(meta, f, indent=1)
    if len(raw) <= 4 << 20:
        with gzip.open(os.path.join(gold, run.name + ".pileup.bin.gz"), "wb", compresslevel=9) as f:
            f.write(raw)
    with gzip.open(os.path.join(gold, run.name + ".indel.norm.txt.gz"), "wt", compresslevel=9) as f:
        f.write(normalise_indel(os.path.join(work, run.name + ".indel.txt.gz")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--work", default="/tmp/golden")
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--hash-idx", action="store_true", help="also sha256 the inflated 16 GiB .idx (minutes)")
    ap.add_argument("fixtures", nargs="+")
    a = ap.parse_args()
    for name in a.fixtures:
        print("fixture", name, flush=True)
        fx = fixtures_def.FIXTURES[name]()
        work = synth.ensure_dir(os.path.join(a.work, name))
        gold = synth.ensure_dir(os.path.join(ROOT, "tests", "golden", name))
        build_index(work, fx)
        meta = {"sdx": open(os.path.join(work, "g.sdx")).read(),
                "mdx_sha256": sha_stream(open, os.path.join(work, "g.mdx"))[0],
                "mdx_bytes": os.path.getsize(os.path.join(work, "g.mdx")),
                "seq_sha256": sha_stream(gzip.open, os.path.join(work, "g.seq"))[0]}
        if a.hash_idx:
            meta["idx_sha256"], meta["idx_bytes"] = sha_stream(gzip.open, os.path.join(work, "g.idx"))
        with open(os.path.join(gold, "index.json"), "w") as f:
            json.dump(meta, f, indent=1)
        for run in fx.runs:
            cmd = run_mapper(work, fx, run, a.threads)
            collect(work, gold, fx, run, cmd)


if __name__ == "__main__":
    main()
