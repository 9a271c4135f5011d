#!/usr/bin/env python
"""Golden files for the pemapper_tsw form: the UNMODIFIED reference pemapper_tsw (oracle/_ref) on the
`tiny` fixture, array mode with per-sample output names and read trimming.  Needs the reference index of `tiny`
in <work>/tiny (tools/make_golden.py tiny).  Writes tests/golden/tsw/.
Layout of a run `r`:  <r>.files.json = the inputs (which reads go to which fastq, sample names, trims, argv),
<sample>.pileup.bin.gz / .summary.txt / .indel.norm.txt.gz per sample, <fastq>.mfile.gz per input file."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from pecaller_b200 import synth  # noqa: E402
import fixtures_def  # noqa: E402
from make_golden import normalise_indel  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def tsw_runs(fx):
    """The two tsw runs, shared with tests/test_cli_files.py."""
    se, pe = fx.runs[0], fx.runs[1]
    return [
        {"name": "tsw_pa", "mode": "pa", "trim": (3, 2), "min_align": 0.85, "max_dist": 500, "min_dist": 0,
         "files": [("pa_A_1.fq", "pa_A_2.fq", "sampA", (0, 600)), ("pa_B_1.fq", "pa_B_2.fq", "sampB", (600, 1000))],
         "reads": (pe.reads1, pe.reads2)},
        {"name": "tsw_sa", "mode": "sa", "trim": (0, 4), "min_align": 0.9,
         "files": [("sa_1.fq", None, "s1", (0, 700)), ("sa_2.fq", None, "s1", (700, 1200)), ("sa_3.fq", None, "s2", (1200, 2000))],
         "reads": (se.reads1, None)},
    ]


def main():
    work = os.path.join(sys.argv[1] if len(sys.argv) > 1 else "/tmp/golden", "tiny")
    gold = synth.ensure_dir(os.path.join(ROOT, "tests", "golden", "tsw"))
    fx = fixtures_def.FIXTURES["tiny"]()
    for run in tsw_runs(fx):
        r1, r2 = run["reads"]
        with open(os.path.join(work, run["name"] + ".arr1"), "w") as a1:
            for f1, f2, samp, (lo, hi) in run["files"]:
                synth.write_fastq(os.path.join(work, f1), r1[lo:hi])
                a1.write("%s\t%s\n" % (f1, samp))
        if r2 is not None:
            with open(os.path.join(work, run["name"] + ".arr2"), "w") as a2:
                for f1, f2, samp, (lo, hi) in run["files"]:
                    synth.write_fastq(os.path.join(work, f2), r2[lo:hi])
                    a2.write("%s\n" % f2)
        n = max(hi - lo for *_, (lo, hi) in run["files"])
        t0, t1 = run["trim"]
        if r2 is not None:
            cmd = ["pemapper_tsw", "unused_base", "g.sdx", "pa", run["name"] + ".arr1", run["name"] + ".arr2", str(run["max_dist"]),
                   str(run["min_dist"]), "n", repr(run["min_align"]), "8", str(n + 8), str(t0), str(t1)]
        else:
            cmd = ["pemapper_tsw", "unused_base", "g.sdx", "sa", run["name"] + ".arr1", "n", repr(run["min_align"]), "8", str(n + 8),
                   str(t0), str(t1)]
        subprocess.run([os.path.join(REF, cmd[0])] + cmd[1:], cwd=work, stdout=subprocess.DEVNULL, check=True)
        print(run["name"], " ".join(cmd), flush=True)
        for samp in sorted({f[2] for f in run["files"]}):
            raw = gzip.open(os.path.join(work, samp + ".pileup.gz"), "rb").read()
            with gzip.open(os.path.join(gold, "%s.%s.pileup.bin.gz" % (run["name"], samp)), "wb", compresslevel=9) as f:
                f.write(raw)
            open(os.path.join(gold, "%s.%s.summary.txt" % (run["name"], samp)), "w").write(
                open(os.path.join(work, samp + ".summary.txt")).read())
            with gzip.open(os.path.join(gold, "%s.%s.indel.norm.txt.gz" % (run["name"], samp)), "wt", compresslevel=9) as f:
                f.write(normalise_indel(os.path.join(work, samp + ".indel.txt.gz")))
            print("  sample", samp, len(raw) // 16, "records")
        for f1, f2, samp, (lo, hi) in run["files"]:
            for fq in (f1, f2):
                if fq:
                    raw = np.fromfile(os.path.join(work, fq + ".mfile"), dtype=np.uint32)
                    assert raw.shape[0] == hi - lo
                    with gzip.open(os.path.join(gold, "%s.%s.mfile.gz" % (run["name"], fq)), "wb", compresslevel=9) as f:
                        f.write(raw.tobytes())
        json.dump({"argv": cmd[1:]}, open(os.path.join(gold, run["name"] + ".json"), "w"))


if __name__ == "__main__":
    main()
