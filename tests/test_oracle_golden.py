"""CPU: the oracle (oracle/pemap_oracle.c) against the committed outputs of the UNMODIFIED reference binaries
(tests/golden, produced by tools/make_golden.py from /root/reference).  This is what pins the oracle."""
import hashlib

import numpy as np
import pytest

import golden_io as gio
import oracle_lib as ol

FAST = ["tiny", "edge9", "bis"]
ALL = ["tiny", "edge9", "cfg1", "pe150", "repeat"]


def _check_fixture(fx):
    assert gio.have(fx.name), "golden vectors for %s missing (tools/make_golden.py)" % fx.name
    bis = int(getattr(fx, "bisulfite", False))
    o = ol.Oracle(fx.genome, ol.default_params(is_bisulfite=bis))
    meta = gio.index_meta(fx.name)
    # index_genome_whole parity: .mdx byte for byte, .sdx text
    assert hashlib.sha256(o.mers().tobytes()).hexdigest() == meta["mdx_sha256"]
    cs = o.contig_starts()
    sdx = "%d\n" % len(fx.genome) + "".join("%d\t%s\n" % (cs[i + 1] - cs[i], fx.names[i]) for i in range(len(fx.genome))) + "16\n"
    assert sdx == meta["sdx"]
    if "idx_sha256" in meta and o.genome_size <= 100_000:
        pass  # the 16 GiB .idx stream is checked by tools/pin_index.py (minutes), not here
    for run in fx.runs:
        o.reset()
        o.set_params(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist,
                     is_bisulfite=bis)
        m1, m2, ty = o.map_batch(run.reads1, run.reads2, nthreads=8)
        g1 = gio.mfile(fx.name, run.name, 1)
        assert np.array_equal(g1, m1), "%s/%s: .mfile of mate 1 differs" % (fx.name, run.name)
        if run.paired:
            assert np.array_equal(gio.mfile(fx.name, run.name, 2), m2), "%s/%s: .mfile of mate 2 differs" % (fx.name, run.name)
        rec = o.records()
        pm = gio.pileup_meta(fx.name, run.name)
        assert rec.shape[0] == pm["n_records"]
        assert hashlib.sha256(rec.tobytes()).hexdigest() == pm["sha256"], "%s/%s: pileup bytes differ" % (fx.name, run.name)
        counts, head = gio.summary(fx.name, run.name)
        names = gio.PAIR_NAMES if run.paired else gio.SINGLE_NAMES
        got = np.bincount(ty, minlength=9)
        for code, nm in names.items():
            assert counts[nm] == got[code], (fx.name, run.name, nm)
        assert got.sum() == counts["All"]
        # insertion strings per site as a multiset (order inside a site is thread-dependent in the reference)
        gold_ins = sorted((p[1], s) for p in gio.indel_lines(fx.name, run.name) for s in p[7])
        # oracle positions are concatenated coordinates; compare the multiset of strings and the per-site counts
        oins = o.insertions()
        assert sorted(s for _, s in oins) == sorted(s for _, s in gold_ins)
        assert int(rec["c"][:, 5].sum()) == len(oins)
    o.close()


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_outputs(name, get_fixture, oracle_built):
    _check_fixture(get_fixture(name))


@pytest.mark.slow
@pytest.mark.parametrize("name", [n for n in ALL if n not in FAST])
def test_oracle_matches_reference_outputs_large(name, get_fixture, oracle_built):
    if not gio.have(name):
        pytest.skip("golden for %s not generated" % name)
    _check_fixture(get_fixture(name))


def test_oracle_stage_probes(get_fixture, oracle_built):
    """initial_map / window / sw_align agree with what the batch path did for the same read."""
    fx = get_fixture("tiny")
    o = ol.Oracle(fx.genome)
    run = fx.runs[0]
    m1, m2, ty, det = o.map_batch(run.reads1[:200], None, detail=True)
    for i in range(200):
        spots, orients = o.initial_map(run.reads1[i].tobytes())
        assert spots.shape[0] == det["hits1"][i]
        if det["best1"][i] >= 0:
            b = det["best1"][i]
            ch, start, blen = o.window(int(spots[b]), 100)
            seq = run.reads1[i].tobytes()
            if orients[b]:
                from pecaller_b200.synth import revcomp_rows
                seq = revcomp_rows(run.reads1[i:i + 1])[0].tobytes()
            sc, st = o.sw_align(start, blen, seq)
            assert sc == det["score1"][i]
            assert start + st[1] + 1 == m1[i]
    o.close()
