"""GPU: the C host (pecaller_b200/host/pemapper_gpu: reference argv, FASTQ reader and writers over the C-ABI) against
the committed output FILES of the unmodified reference pemapper (tests/golden, tools/make_golden.py):
.mfile and the inflated .pileup.gz byte for byte, .summary.txt verbatim, .indel.txt.gz up to the order of the
insertion strings of a site (not deterministic in the reference either)."""
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

import golden_io as gio
import oracle_lib as ol
from pecaller_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "pecaller_b200", "host", "pemapper_gpu")


def _normalise_indel(path):
    lines = []
    with gzip.open(path, "rt") as f:
        for ln in f.read().split("\n"):
            parts = ln.split("\t")
            if len(parts) > 7 and parts[0] != "Fragment":
                parts = parts[:7] + sorted(parts[7:])
            lines.append("\t".join(parts))
    return "\n".join(lines)


@pytest.mark.parametrize("name,gpus", [("tiny", 1), ("edge9", 1), ("bis", 1), ("edge9", 2)])
def test_cli_files_match_reference(name, gpus, get_fixture, oracle_built, tmp_path):
    """gpus == 2: PEMAP_GPUS=2 with small batches - one submitting thread per GPU, batches alternate between the GPUs,
    counters summed over NVLink peer memory (pemap_reduce_counts_peer); the files must not change."""
    if not os.path.exists(CLI):
        pytest.fail("C host not built: run __graft_entry__.build()")
    if gpus > 1:
        import torch
        if torch.cuda.device_count() < gpus:
            pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (gpus, gpus))
    fx = get_fixture(name)
    work = str(tmp_path)
    oracle = ol.Oracle(fx.genome)
    oracle.write_index(os.path.join(work, "g"), fx.names, with_idx=False)   # .sdx/.seq (+.mdx); .idx is rebuilt on the GPU
    oracle.close()
    assert open(os.path.join(work, "g.sdx")).read() == gio.index_meta(name)["sdx"]
    env = dict(os.environ, PEMAP_DEVICE_INDEX="1")
    if gpus > 1:
        env.update(PEMAP_GPUS=str(gpus), PEMAP_BATCH="777")
    checked = 0
    for run in fx.runs:
        if not gio.have(name, run.name):
            continue
        n = run.reads1.shape[0]
        synth.write_fastq(os.path.join(work, run.name + "_1.fq"), run.reads1)
        bis = "y" if run.bisulfite else "n"
        if run.paired:
            synth.write_fastq(os.path.join(work, run.name + "_2.fq"), run.reads2)
            cmd = [CLI, run.name, "g.sdx", "p", run.name + "_1.fq", run.name + "_2.fq", str(run.max_dist), str(run.min_dist),
                   bis, repr(run.min_align), "4", str(n + 8)]
        else:
            cmd = [CLI, run.name, "g.sdx", "s", run.name + "_1.fq", bis, repr(run.min_align), "4", str(n + 8)]
        r = subprocess.run(cmd, cwd=work, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        tag = "%s/%s" % (name, run.name)
        m1 = np.fromfile(os.path.join(work, run.name + "_1.fq.mfile"), dtype=np.uint32)
        assert np.array_equal(m1, gio.mfile(name, run.name, 1)), tag + ": .mfile 1"
        if run.paired:
            m2 = np.fromfile(os.path.join(work, run.name + "_2.fq.mfile"), dtype=np.uint32)
            assert np.array_equal(m2, gio.mfile(name, run.name, 2)), tag + ": .mfile 2"
        raw = gzip.open(os.path.join(work, run.name + ".pileup.gz"), "rb").read()
        assert hashlib.sha256(raw).hexdigest() == gio.pileup_meta(name, run.name)["sha256"], tag + ": .pileup.gz"
        gold_summary = open(os.path.join(gio.GOLD, name, run.name + ".summary.txt")).read()
        assert open(os.path.join(work, run.name + ".summary.txt")).read() == gold_summary, tag + ": .summary.txt"
        gold_indel = gzip.open(os.path.join(gio.GOLD, name, run.name + ".indel.norm.txt.gz"), "rt").read()
        assert _normalise_indel(os.path.join(work, run.name + ".indel.txt.gz")) == gold_indel, tag + ": .indel.txt.gz"
        checked += 1
    assert checked > 0


def test_cli_tsw_form_matches_reference(get_fixture, oracle_built, tmp_path):
    """pemapper_tsw form: array files with a sample column, read trimming, outputs written and counters
    zeroed whenever the sample changes - against the files of the unmodified reference pemapper_tsw
    (tests/golden/tsw, tools/make_golden_tsw.py)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_tsw import tsw_runs
    gold = os.path.join(gio.GOLD, "tsw")
    fx = get_fixture("tiny")
    work = str(tmp_path)
    oracle = ol.Oracle(fx.genome)
    oracle.write_index(os.path.join(work, "g"), fx.names, with_idx=False)
    oracle.close()
    env = dict(os.environ, PEMAP_DEVICE_INDEX="1", PEMAP_BATCH="333")
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
        import json
        argv = json.load(open(os.path.join(gold, run["name"] + ".json")))["argv"]
        r = subprocess.run([CLI] + argv, cwd=work, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        for samp in sorted({f[2] for f in run["files"]}):
            tag = "%s/%s" % (run["name"], samp)
            raw = gzip.open(os.path.join(work, samp + ".pileup.gz"), "rb").read()
            want = gzip.open(os.path.join(gold, "%s.%s.pileup.bin.gz" % (run["name"], samp)), "rb").read()
            assert raw == want, tag + ": .pileup.gz"
            assert open(os.path.join(work, samp + ".summary.txt")).read() == \
                open(os.path.join(gold, "%s.%s.summary.txt" % (run["name"], samp))).read(), tag + ": .summary.txt"
            assert _normalise_indel(os.path.join(work, samp + ".indel.txt.gz")) == \
                gzip.open(os.path.join(gold, "%s.%s.indel.norm.txt.gz" % (run["name"], samp)), "rt").read(), tag + ": indel"
        for f1, f2, samp, (lo, hi) in run["files"]:
            for fq in (f1, f2):
                if fq:
                    got = np.fromfile(os.path.join(work, fq + ".mfile"), dtype=np.uint32)
                    want = np.frombuffer(gzip.open(os.path.join(gold, "%s.%s.mfile.gz" % (run["name"], fq))).read(), dtype=np.uint32)
                    assert np.array_equal(got, want), "%s: %s.mfile" % (run["name"], fq)


def test_index_genome_gpu_files_match_reference(get_fixture, tmp_path):
    """index_genome_gpu (device-built index written as .sdx/.seq/.idx/.mdx) against the files of the unmodified
    index_genome_whole: .sdx text, sha256 of .mdx and of the inflated .seq and .idx (16 GiB stream) - SURVEY 8f-2."""
    tool = os.path.join(ROOT, "pecaller_b200", "host", "index_genome_gpu")
    if not os.path.exists(tool):
        pytest.fail("index_genome_gpu not built: run __graft_entry__.build()")
    for name, full_idx in (("edge9", False), ("tiny", True)):
        fx = get_fixture(name)
        meta = gio.index_meta(name)
        work = str(tmp_path / name)
        os.makedirs(work)
        synth.write_fasta(os.path.join(work, "g.fa"), fx.genome, fx.names)
        env = dict(os.environ, PEMAP_INDEX_SKIP_IDX="0" if full_idx else "1")
        r = subprocess.run([tool, "g.fa", "g", "n"], cwd=work, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert open(os.path.join(work, "g.sdx")).read() == meta["sdx"], name + ": .sdx"
        assert hashlib.sha256(open(os.path.join(work, "g.mdx"), "rb").read()).hexdigest() == meta["mdx_sha256"], name + ": .mdx"
        assert hashlib.sha256(gzip.open(os.path.join(work, "g.seq"), "rb").read()).hexdigest() == meta["seq_sha256"], name + ": .seq"
        if full_idx and "idx_sha256" in meta:
            h = hashlib.sha256()
            n = 0
            with gzip.open(os.path.join(work, "g.idx"), "rb") as f:
                while True:
                    b = f.read(1 << 26)
                    if not b:
                        break
                    h.update(b)
                    n += len(b)
            assert n == meta["idx_bytes"] and h.hexdigest() == meta["idx_sha256"], name + ": inflated .idx"
            # the drop-in path: pemapper_gpu WITHOUT PEMAP_DEVICE_INDEX loads these .idx / .mdx files as the reference's
            # init_index_buffer does (pemapper.c:2129-2155) and hands the host arrays to pemap_init
            run = next(r for r in fx.runs if gio.have(name, r.name) and not r.paired)
            synth.write_fastq(os.path.join(work, "q.fq"), run.reads1)
            env2 = {k: v for k, v in os.environ.items() if k != "PEMAP_DEVICE_INDEX"}
            r = subprocess.run([CLI, "q", "g.sdx", "s", "q.fq", "y" if run.bisulfite else "n", repr(run.min_align), "4",
                                str(run.reads1.shape[0] + 8)], cwd=work, env=env2, capture_output=True, text=True, timeout=900)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            assert np.array_equal(np.fromfile(os.path.join(work, "q.fq.mfile"), dtype=np.uint32), gio.mfile(name, run.name, 1))
            raw = gzip.open(os.path.join(work, "q.pileup.gz"), "rb").read()
            assert hashlib.sha256(raw).hexdigest() == gio.pileup_meta(name, run.name)["sha256"], "pileup through pemap_init"
