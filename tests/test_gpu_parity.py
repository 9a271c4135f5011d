"""GPU: the CUDA path, called through the C-ABI, against the oracle on the same seeded inputs (every stage:
index, candidate lists, strand, score bits, m1/m2, mapping type, pileup records, insertion strings) and against
the committed outputs of the reference binaries (tests/golden)."""
import hashlib

import numpy as np
import pytest

import golden_io as gio
import oracle_lib as ol
import pecaller_b200 as pb

pytestmark = pytest.mark.gpu


def _run_both(mapper, oracle, run, n=None, threads=16):
    r1 = run.reads1[:n]
    r2 = run.reads2[:n] if run.paired else None
    kw = dict(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist,
              is_bisulfite=int(run.bisulfite))
    oracle.reset()
    oracle.set_params(**kw)
    mapper.reset_counts()
    mapper.set_params(**kw)
    mapper.keep(pb.KEEP_DETAIL | pb.KEEP_CANDIDATES)
    o = oracle.map_batch(r1, r2, nthreads=threads, detail=True)
    g = mapper.map_batch(r1, r2)
    return r1, r2, o, g


def _assert_parity(name, mapper, oracle, run, n=None, cand_step=37):
    r1, r2, (om1, om2, oty, odet), (gm1, gm2, gty) = _run_both(mapper, oracle, run, n)
    tag = "%s/%s" % (name, run.name)
    assert np.array_equal(om1, gm1), tag + ": m1"
    assert np.array_equal(om2, gm2), tag + ": m2"
    assert np.array_equal(oty, gty), tag + ": mapping_type"
    gdet = mapper.detail(r1.shape[0])
    for f in ("hits1", "hits2", "best1", "best2", "orient1", "orient2"):
        assert np.array_equal(odet[f], gdet[f]), tag + ": " + f
    for f in ("score1", "score2"):  # bit pattern of the reference's double
        assert np.array_equal(odet[f].view(np.uint64), gdet[f].view(np.uint64)), tag + ": " + f
    for i in range(0, r1.shape[0], cand_step):
        os_, oo = oracle.initial_map(r1[i].tobytes())
        gs, go = mapper.candidates(i, 0)
        assert np.array_equal(os_, gs) and np.array_equal(oo, go), tag + ": candidates of read %d" % i
    orec = oracle.records()
    grec, gins = mapper.finish()
    assert orec.tobytes() == grec.tobytes(), tag + ": pileup records"
    oins = oracle.insertions()
    assert oins == sorted(gins), tag + ": insertion strings"
    # default mode: integer DPX scoring + fp64 replay of rational ties must give the same bytes
    # (keep(DETAIL) above forced the all-fp64 path, whose scores are the reference's doubles bit for bit)
    mapper.keep(0)
    mapper.reset_counts()
    mapper.reset_stats()
    fm1, fm2, fty = mapper.map_batch(r1, r2)
    assert np.array_equal(om1, fm1), tag + ": m1 (integer path)"
    assert np.array_equal(om2, fm2), tag + ": m2 (integer path)"
    assert np.array_equal(oty, fty), tag + ": mapping_type (integer path)"
    frec, fins = mapper.finish()
    assert orec.tobytes() == frec.tobytes(), tag + ": pileup records (integer path)"
    assert oins == sorted(fins), tag + ": insertion strings (integer path)"
    st = mapper.stats()
    print("%s: %d read-mates, %d replayed in fp64 (%.2f%%)" % (tag, st["reads"], st["replayed"], 100.0 * st["replayed"] / max(1, st["reads"])))
    return gm1, gm2, gty, grec


@pytest.fixture(scope="module")
def ctx_cache():
    cache = {}
    yield cache
    for m, o in cache.values():
        m.close()
        o.close()


def _ctx(cache, fx):
    if fx.name not in cache:
        for name in list(cache):  # one index (16 GiB pos_index + chunk buffers, ~35 GB of HBM) at a time
            m, o = cache.pop(name)
            m.close()
            o.close()
        bis = int(getattr(fx, "bisulfite", False))  # baked into the index (index_genome_whole.c:174-175)
        cache[fx.name] = (pb.PEMapper.from_genome(fx.genome, pb.default_params(is_bisulfite=bis)),
                          ol.Oracle(fx.genome, ol.default_params(is_bisulfite=bis)))
    return cache[fx.name]


@pytest.mark.parametrize("name", ["tiny", "edge9", "cfg1", "pe150", "repeat", "bis"])
def test_cuda_matches_oracle_and_reference(name, get_fixture, ctx_cache, oracle_built):
    fx = get_fixture(name)
    mapper, oracle = _ctx(ctx_cache, fx)
    # index_genome_whole parity of the device index builder
    assert np.array_equal(mapper.read_mers(), oracle.mers())
    rng = np.random.default_rng(1)
    for w in rng.integers(0, 1 << 32, size=200, dtype=np.uint64):
        assert int(mapper.read_pos_index(int(w), 1)[0]) == oracle.pos_index(int(w))
    assert int(mapper.read_pos_index(1 << 32, 1)[0]) == oracle.mers().shape[0]
    for run in fx.runs:
        gm1, gm2, gty, grec = _assert_parity(name, mapper, oracle, run)
        if gio.have(name, run.name):  # the reference binary's own files
            assert np.array_equal(gio.mfile(name, run.name, 1), gm1)
            if run.paired:
                assert np.array_equal(gio.mfile(name, run.name, 2), gm2)
            pmeta = gio.pileup_meta(name, run.name)
            assert hashlib.sha256(grec.tobytes()).hexdigest() == pmeta["sha256"]
            counts, _ = gio.summary(name, run.name)
            names = gio.PAIR_NAMES if run.paired else gio.SINGLE_NAMES
            got = np.bincount(gty, minlength=9)
            for code, nm in names.items():
                assert counts[nm] == got[code]


@pytest.mark.parametrize("name", ["tiny", "bis"])
def test_device_index_equals_reference_idx_stream(name, get_fixture):
    """The device index builder against the reference index_genome_whole's FILES: sha256 of the whole pos_index
    (2^32+1 words = the inflated .idx, 16 GiB) and of mers (= .mdx), normal and bisulfite index."""
    meta = gio.index_meta(name)
    if "idx_sha256" not in meta:
        pytest.skip("reference .idx stream not hashed for %s" % name)
    fx = get_fixture(name)
    mapper = pb.PEMapper.from_genome(fx.genome, pb.default_params(is_bisulfite=int(getattr(fx, "bisulfite", False))))
    assert hashlib.sha256(mapper.read_mers().tobytes()).hexdigest() == meta["mdx_sha256"]
    h = hashlib.sha256()
    total = (1 << 32) + 1
    step = 1 << 28
    for first in range(0, total, step):
        h.update(mapper.read_pos_index(first, min(step, total - first)).tobytes())
    mapper.close()
    assert h.hexdigest() == meta["idx_sha256"]


def test_pointer_entry_and_chunking(get_fixture, ctx_cache, oracle_built, monkeypatch):
    """pemap_map_batch (char** form of PTHREAD_DATA_NODE) == pemap_map_batch_rows; results independent of batching."""
    fx = get_fixture("tiny")
    mapper, oracle = _ctx(ctx_cache, fx)
    run = fx.runs[1]
    kw = dict(min_align=run.min_align, pair_flag=1, min_dist=run.min_dist, max_dist=run.max_dist)
    mapper.set_params(**kw)
    mapper.reset_counts()
    a = mapper.map_batch(run.reads1, run.reads2)
    rec_a, ins_a = mapper.finish()
    mapper.reset_counts()
    l1 = [r.tobytes() for r in run.reads1]
    l2 = [r.tobytes() for r in run.reads2]
    parts = [mapper.map_pointers(l1[s:s + 333], l2[s:s + 333]) for s in range(0, len(l1), 333)]
    b = [np.concatenate([p[k] for p in parts]) for k in range(3)]
    rec_b, ins_b = mapper.finish()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert rec_a.tobytes() == rec_b.tobytes()
    assert sorted(ins_a) == sorted(ins_b)


def test_edge_inputs(get_fixture, ctx_cache, oracle_built):
    """Empty batch, all-N reads, poly-T reads (k-mer 0xFFFFFFFF quirk), mixed lengths, lower-case, counter reset."""
    fx = get_fixture("tiny")
    mapper, oracle = _ctx(ctx_cache, fx)
    mapper.set_params(min_align=0.9, pair_flag=0)
    oracle.set_params(min_align=0.9, pair_flag=0)
    m1, m2, ty = mapper.map_batch(np.zeros((0, 100), dtype=np.uint8))
    assert m1.shape == (0,)
    g = fx.genome[0]
    rows = []
    rows.append(np.full(100, ord("N"), np.uint8))
    rows.append(np.full(100, ord("T"), np.uint8))
    rows.append(np.full(100, ord("A"), np.uint8))
    r = g[1000:1100].copy(); r[::9] = ord("N"); rows.append(r)            # 12 N >= 1+len/10 -> filtered
    r = g[2000:2100].copy(); r[5] = ord("N"); r[50] = ord("n"); rows.append(r)
    r = g[3000:3100].copy(); r[10:14] = np.frombuffer(b"acgt", np.uint8); rows.append(r)   # lower case
    r = g[4000:4100].copy(); r[30] = ord("R"); r[31] = ord("Y"); rows.append(r)           # IUPAC codes
    reads = np.stack(rows)
    oracle.reset(); mapper.reset_counts()
    o = oracle.map_batch(reads)
    gq = mapper.map_batch(reads)
    for x, y in zip(o, gq):
        assert np.array_equal(x, y)
    assert oracle.records().tobytes() == mapper.finish()[0].tobytes()
    # ragged lengths in one batch through the rows entry
    lens = np.array([100, 64, 17, 16, 150, 33, 250, 99], dtype=np.int32)
    buf = np.zeros((8, 256), dtype=np.uint8)
    for i, L in enumerate(lens):
        buf[i, :L] = g[5000 + 300 * i: 5000 + 300 * i + L]
    oracle.reset(); mapper.reset_counts()
    om = np.zeros((3, 8), dtype=np.int64)
    for i, L in enumerate(lens):
        a, b, c = oracle.map_batch(buf[i:i + 1, :L])
        om[:, i] = a[0], b[0], c[0]
    gm1, gm2, gty = mapper.map_rows(buf, lens)
    assert np.array_equal(om[0], gm1) and np.array_equal(om[2], gty)
    assert oracle.records().tobytes() == mapper.finish()[0].tobytes()
    mapper.reset_counts()
    assert mapper.finish()[0].shape[0] == 0


def test_traceback_band_handover(get_fixture, oracle_built, monkeypatch):
    """The traceback kernels keep decision bits only for a band around the winner's end diagonal; walks that leave it
    are redone by the kernel with the full store.  PEMAP_BAND_HALF=0 narrows the band to one lane so that every gapped
    read of pe150 takes that hand-over; the pileup must not change."""
    fx = get_fixture("pe150")
    run = fx.runs[0]
    oracle = ol.Oracle(fx.genome)
    kw = dict(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist)
    oracle.set_params(**kw)
    n = 20000
    o = oracle.map_batch(run.reads1[:n], run.reads2[:n] if run.paired else None, nthreads=16)
    orec, oins = oracle.records(), oracle.insertions()
    for half in ("0", "1"):
        monkeypatch.setenv("PEMAP_BAND_HALF", half)
        mapper = pb.PEMapper.from_genome(fx.genome)
        mapper.set_params(**kw)
        g = mapper.map_batch(run.reads1[:n], run.reads2[:n] if run.paired else None)
        for x, y in zip(o[:3], g):
            assert np.array_equal(x, y)
        rec, ins = mapper.finish()
        assert orec.tobytes() == rec.tobytes(), "band half %s: pileup records" % half
        assert oins == sorted(ins)
        mapper.close()
    oracle.close()


@pytest.mark.parametrize("name", ["pe150", "repeat", "bis"])
def test_diagonal_certificate_changes_nothing(name, get_fixture, monkeypatch):
    """k_diag_certify settles candidates whose result follows from their ungapped diagonals (sw_int16.cuh); with
    PEMAP_CERTIFY=0 every candidate goes through the DP kernel instead.  Loci, mapping types, pileup records and
    insertions must be identical, and the certificate must actually have fired."""
    fx = get_fixture(name)
    bis = int(getattr(fx, "bisulfite", False))  # baked into the index (index_genome_whole.c:174-175)
    for run in fx.runs:
        kw = dict(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist,
                  is_bisulfite=int(run.bisulfite))
        n = min(20000, run.reads1.shape[0])
        out = {}
        for cert in ("0", "1"):
            monkeypatch.setenv("PEMAP_CERTIFY", cert)
            mapper = pb.PEMapper.from_genome(fx.genome, pb.default_params(is_bisulfite=bis))
            mapper.set_params(**kw)
            g = mapper.map_batch(run.reads1[:n], run.reads2[:n] if run.paired else None)
            rec, ins = mapper.finish()
            st = mapper.stats()
            out[cert] = (g, rec.tobytes(), sorted(ins), st["sw_cells_certified"])
            mapper.close()
        for x, y in zip(out["0"][0], out["1"][0]):
            assert np.array_equal(x, y), "%s/%s: per-read results differ with the certificate on" % (name, run.name)
        assert out["0"][1] == out["1"][1], "%s/%s: pileup records" % (name, run.name)
        assert out["0"][2] == out["1"][2], "%s/%s: insertions" % (name, run.name)
        assert out["0"][3] == 0
        if name != "repeat":
            assert out["1"][3] > 0, "the certificate never fired"


@pytest.mark.parametrize("name", ["pe150", "repeat", "bis", "edge9"])
def test_seed_index_layouts_agree(name, get_fixture, monkeypatch):
    """The seed stage reads a device-private rotated bucket index (seed_rbi.cuh); PEMAP_SEED=legacy runs the round-1
    kernel over pos_index / mers as the reference lays them out, and PEMAP_RBI_CAP=8 pushes nearly every read-mate
    through the second (global-memory) pass of the new kernel.  Candidate lists in order, loci, types, pileup records
    and insertions must be identical in all three."""
    fx = get_fixture(name)
    bis = int(getattr(fx, "bisulfite", False))
    for run in fx.runs:
        kw = dict(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist,
                  is_bisulfite=int(run.bisulfite))
        n = min(20000, run.reads1.shape[0])
        out = {}
        for mode in ("legacy", "rbi", "rbi-big"):
            monkeypatch.setenv("PEMAP_SEED", "legacy" if mode == "legacy" else "rbi")
            monkeypatch.setenv("PEMAP_RBI_CAP", "8" if mode == "rbi-big" else "512")
            mapper = pb.PEMapper.from_genome(fx.genome, pb.default_params(is_bisulfite=bis))
            mapper.set_params(**kw)
            mapper.keep(pb.KEEP_CANDIDATES)
            g = mapper.map_batch(run.reads1[:n], run.reads2[:n] if run.paired else None)
            cands = [mapper.candidates(i, m) for i in range(0, n, 7) for m in range(2 if run.paired else 1)]
            rec, ins = mapper.finish()
            st = mapper.stats()
            out[mode] = (g, cands, rec.tobytes(), sorted(ins), st["candidates"], st["mer_positions"])
            mapper.close()
        for mode in ("rbi", "rbi-big"):
            tag = "%s/%s %s vs legacy" % (name, run.name, mode)
            for x, y in zip(out["legacy"][0], out[mode][0]):
                assert np.array_equal(x, y), tag + ": per-read results"
            for (s0, o0), (s1, o1) in zip(out["legacy"][1], out[mode][1]):
                assert np.array_equal(s0, s1) and np.array_equal(o0, o1), tag + ": candidate lists"
            assert out["legacy"][2] == out[mode][2], tag + ": pileup records"
            assert out["legacy"][3] == out[mode][3], tag + ": insertions"
            assert out["legacy"][4] == out[mode][4], tag + ": candidate totals"


def test_bounded_finish_and_insertion_spill(get_fixture, oracle_built, monkeypatch):
    """pemap_finish_stream compacts the counters through bounded windows and the insertion append buffer spills to the
    host when a chunk leaves it more than half full: with 4096-site windows, 32 KB of insertion buffer (a 1024-pair
    chunk appends ~4 KB; the fill level is seen two chunks late because chunks are pipelined) and 1024-read chunks the records, the callback's windows and the insertion multiset equal the oracle's."""
    fx = get_fixture("pe150")
    run = fx.runs[0]
    n = 20000
    kw = dict(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist)
    oracle = ol.Oracle(fx.genome)
    oracle.set_params(**kw)
    o = oracle.map_batch(run.reads1[:n], run.reads2[:n] if run.paired else None, nthreads=16)
    orec, oins = oracle.records(), oracle.insertions()
    assert len(oins) > 200
    monkeypatch.setenv("PEMAP_FINISH_SITES", "4096")
    monkeypatch.setenv("PEMAP_INS_BYTES", "32768")
    monkeypatch.setenv("PEMAP_CHUNK", "1024")
    mapper = pb.PEMapper.from_genome(fx.genome)
    mapper.set_params(**kw)
    g = mapper.map_batch(run.reads1[:n], run.reads2[:n] if run.paired else None)
    for x, y in zip(o[:3], g):
        assert np.array_equal(x, y)
    parts = []
    total = mapper.finish_stream(lambda r: parts.append(r.copy()))
    assert total == orec.shape[0] and len(parts) > 10
    assert np.concatenate(parts).tobytes() == orec.tobytes()
    assert sorted(mapper.insertions()) == oins
    rec, ins = mapper.finish()               # the collecting wrapper, again (counters are not cleared)
    assert rec.tobytes() == orec.tobytes() and sorted(ins) == oins
    mapper.close()
    # an append buffer too small for ONE chunk fails that batch at once instead of at the end of the run
    monkeypatch.setenv("PEMAP_INS_BYTES", "1024")
    monkeypatch.setenv("PEMAP_CHUNK", "16384")
    mapper = pb.PEMapper.from_genome(fx.genome)
    mapper.set_params(**kw)
    with pytest.raises(pb.PemapError):
        mapper.map_batch(run.reads1[:n], run.reads2[:n] if run.paired else None)
    mapper.close()
    oracle.close()


def _two_gpu_worker(rank, world, port, out_path):
    import os
    import torch
    import torch.distributed as dist
    import fixtures_def
    from pecaller_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fx = fixtures_def.FIXTURES["pe150"]()
    run = fx.runs[0]
    n = 30000
    kw = dict(min_align=run.min_align, pair_flag=1, min_dist=run.min_dist, max_dist=run.max_dist)
    mapper = pb.PEMapper.from_genome(fx.genome, device=rank)
    mapper.set_params(**kw)
    ranges = sharding.shard_batches(n, rank, world, 4096)
    parts = [mapper.map_batch(run.reads1[a:b], run.reads2[a:b]) for a, b in ranges]
    m1, m2, ty = (np.concatenate([p[k] for p in parts]) for k in range(3))
    # (a) slice-wise sum over NVLink peer memory: every rank ends up with the final counters of its slice and compacts it
    red = sharding.SliceReducer(mapper)
    lo, hi = red.reduce_scatter()
    parts_rec = []
    mapper.finish_stream(lambda r: parts_rec.append(r.copy()), site_range=(lo, hi))
    mine = np.concatenate(parts_rec) if parts_rec else np.zeros(0, dtype=pb.RECORD_DTYPE)
    all_rec = [None] * world
    dist.all_gather_object(all_rec, mine)
    # (b) NCCL sum onto rank 0, chromosome by chromosome (the slices above already hold sums: undo nothing, the reduce
    # below runs on a second mapper's counters)
    mapper2 = pb.PEMapper.from_genome(fx.genome, device=rank)
    mapper2.set_params(**kw)
    for a, b in ranges:
        mapper2.map_batch(run.reads1[a:b], run.reads2[a:b])
    sharding.reduce_counts(sharding.counts_tensor(mapper2, torch.device("cuda", rank)), dst=0)
    torch.cuda.synchronize()
    res = sharding.gather_results(n, ranges, m1, m2, ty, dst=0)
    rec, ins = mapper2.finish()
    mapper2.close()
    all_ins = [None] * world
    dist.all_gather_object(all_ins, ins)
    if rank == 0:
        assert np.concatenate(all_rec).tobytes() == rec.tobytes(), "slice-wise NVLink sum != NCCL reduce"
        np.savez(out_path, rec=rec, m1=res[0], m2=res[1], ty=res[2])
        import pickle
        pickle.dump(sorted(sum(all_ins, [])), open(out_path + ".ins", "wb"))
    mapper.close()
    dist.destroy_process_group()


def test_two_gpus_equal_one(get_fixture, tmp_path):
    """Reads sharded over 2 GPUs + one NCCL sum of the counter arrays == 1 GPU, byte for byte."""
    import pickle
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "two.npz")
    mp.spawn(_two_gpu_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    fx = get_fixture("pe150")
    run = fx.runs[0]
    n = 30000
    mapper = pb.PEMapper.from_genome(fx.genome)
    mapper.set_params(min_align=run.min_align, pair_flag=1, min_dist=run.min_dist, max_dist=run.max_dist)
    m1, m2, ty = mapper.map_batch(run.reads1[:n], run.reads2[:n])
    rec, ins = mapper.finish()
    mapper.close()
    assert np.array_equal(got["m1"], m1) and np.array_equal(got["m2"], m2) and np.array_equal(got["ty"], ty)
    assert got["rec"].tobytes() == rec.tobytes()
    assert pickle.load(open(out + ".ins", "rb")) == sorted(ins)


def test_integer_path_equals_exact_path_at_scale(monkeypatch):
    """Size-independent property at bench scale (cfg2 model, 4 M pairs = 8 M read-mates, ~2.2 M gapped tracebacks):
    the default path (s16x2 scoring, diagonal fast path, integer traceback with tie certificates, fp64 only for
    uncertified ties) must give the same m1/m2/type and byte-identical pileup records and insertion multiset as the
    all-fp64 path (PEMAP_EXACT=1), which evaluates the reference's own double expressions for every candidate and is
    pinned bit for bit to the oracle and the reference binary on the fixtures above."""
    import torch
    import bench
    n = 4_000_000
    dev = torch.device("cuda", 0)
    genome = bench.make_genome(20, 64_000_000)
    gt = torch.from_numpy(genome).to(dev)
    d_r1, d_r2 = bench.torch_reads(gt, n, 77, dev)
    d_len = torch.full((n,), bench.READ_LEN, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    params = pb.default_params(min_align=bench.MIN_ALIGN, pair_flag=1, min_dist=bench.MIN_DIST, max_dist=bench.MAX_DIST)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PEMAP_EXACT", mode)
        mapper = pb.PEMapper.from_genome([genome], params)
        m1 = torch.zeros(n, dtype=torch.int32, device=dev)
        m2 = torch.zeros(n, dtype=torch.int32, device=dev)
        ty = torch.zeros(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        mapper.map_device(n, d_r1.data_ptr(), d_len.data_ptr(), d_r2.data_ptr(), d_len.data_ptr(), bench.STRIDE,
                          bench.READ_LEN, m1.data_ptr(), m2.data_ptr(), ty.data_ptr())
        rec, ins = mapper.finish()
        st = mapper.stats()
        out[mode] = (m1.cpu().numpy(), m2.cpu().numpy(), ty.cpu().numpy(), hashlib.sha256(rec.tobytes()).hexdigest(),
                     rec.shape[0], hashlib.sha256(repr(sorted(ins)).encode()).hexdigest(), len(ins), st)
        mapper.close()
    a, b = out["0"], out["1"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[4] == b[4] and a[3] == b[3], "pileup records differ between the integer and the exact path"
    assert a[6] == b[6] and a[5] == b[5], "insertion strings differ between the integer and the exact path"
    # mass conservation: every counted base / deletion comes from a mapped read-mate, at most len per mate
    mapped = int((a[0] != 0).sum() + (a[1] != 0).sum())
    assert mapped > 0.99 * 2 * n
    assert a[7]["diag_traced"] > 0 and a[7]["exact_traced"] < 0.1 * mapped


def test_peer_reduce_single_process(get_fixture):
    """C-ABI path for single-process multi-GPU hosts: two handles on two GPUs, batches split between them,
    pemap_reduce_counts_peer over NVLink peer memory, then finish on the first == everything on one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    fx = get_fixture("pe150")
    run = fx.runs[0]
    n = 20000
    kw = dict(min_align=run.min_align, pair_flag=1, min_dist=run.min_dist, max_dist=run.max_dist)
    a = pb.PEMapper.from_genome(fx.genome, device=0)
    b = pb.PEMapper.from_genome(fx.genome, device=1)
    one = pb.PEMapper.from_genome(fx.genome, device=0)
    for m in (a, b, one):
        m.set_params(**kw)
    ra = a.map_batch(run.reads1[:n // 2], run.reads2[:n // 2])
    rb = b.map_batch(run.reads1[n // 2:n], run.reads2[n // 2:n])
    r1 = one.map_batch(run.reads1[:n], run.reads2[:n])
    # slice-wise sum first (each GPU pulls its half from the other, then compacts it) ...
    lo_a, hi_a = pb.PEMapper.reduce_scatter_local([a, b], 0)
    lo_b, hi_b = pb.PEMapper.reduce_scatter_local([a, b], 1)
    assert lo_a == 0 and hi_a == lo_b
    sl = []
    a.finish_stream(lambda r: sl.append(r.copy()), site_range=(lo_a, hi_a))
    b.finish_stream(lambda r: sl.append(r.copy()), site_range=(lo_b, hi_b))
    rec_slices = np.concatenate(sl)
    ins_a, ins_b = a.insertions(), b.insertions()
    rec1, ins1 = one.finish()
    for k in range(3):
        assert np.array_equal(np.concatenate([ra[k], rb[k]]), r1[k])
    assert rec_slices.tobytes() == rec1.tobytes()
    assert sorted(ins_a + ins_b) == sorted(ins1)
    # ... and the older whole-array form on fresh counters: b's array added onto a's
    for m in (a, b):
        m.reset_counts()
    a.map_batch(run.reads1[:n // 2], run.reads2[:n // 2])
    b.map_batch(run.reads1[n // 2:n], run.reads2[n // 2:n])
    a.reduce_counts_from(b)
    rec, _ = a.finish()
    assert rec.tobytes() == rec1.tobytes()
    for m in (a, b, one):
        m.close()


@pytest.mark.parametrize("ways", [3, 4, 8])
def test_slice_sum_many_handles_one_gpu(ways, get_fixture, monkeypatch):
    """The N-way slice sum (k_reduce_slice: the unrolled 3- and 7-peer forms and the generic loop) without needing N
    GPUs: N handles on ONE device map N disjoint parts of the batch, every handle pulls its 1/N slice from the others
    and compacts it; slices in rank order == everything mapped through one handle."""
    monkeypatch.setenv("PEMAP_CHUNK", "4096")   # small per-handle scratch: up to nine handles share the device ...
    monkeypatch.setenv("PEMAP_KEEP_INDEX", "0")  # ... and none keeps the 16 GiB file-format pos_index beside its bucket index
    fx = get_fixture("pe150")
    run = fx.runs[0]
    n = 16000
    kw = dict(min_align=run.min_align, pair_flag=1, min_dist=run.min_dist, max_dist=run.max_dist)
    hs = [pb.PEMapper.from_genome(fx.genome, device=0) for _ in range(ways)]
    one = pb.PEMapper.from_genome(fx.genome, device=0)
    for m in hs + [one]:
        m.set_params(**kw)
    cuts = np.linspace(0, n, ways + 1).astype(int)
    for r, m in enumerate(hs):
        m.map_batch(run.reads1[cuts[r]:cuts[r + 1]], run.reads2[cuts[r]:cuts[r + 1]])
    one.map_batch(run.reads1[:n], run.reads2[:n])
    bounds = [pb.PEMapper.reduce_scatter_local(hs, r) for r in range(ways)]
    assert bounds[0][0] == 0 and all(bounds[r][1] == bounds[r + 1][0] for r in range(ways - 1))
    sl = []
    for r, m in enumerate(hs):
        m.finish_stream(lambda x: sl.append(x.copy()), site_range=bounds[r])
    rec1, _ = one.finish()
    assert np.concatenate(sl).tobytes() == rec1.tobytes()
    for m in hs + [one]:
        m.close()


@pytest.mark.parametrize("name", ["pe150", "bis", "edge9"])
def test_packed_reads_equal_ascii_rows(name, get_fixture):
    """pemap_map_batch_packed (2-bit codes + N mask, 64 bytes per 150-bp read; the seed kernel cuts its k-mers out of the
    packed words, the DP kernels read rows unpacked on the device) against pemap_map_batch_rows on the same reads:
    candidates in order, loci, types, pileup records, insertions.  Reads with N (below and above the N filter), ragged
    lengths inside one packed batch and the bisulfite read conversion are included."""
    fx = get_fixture(name)
    bis = int(getattr(fx, "bisulfite", False))
    rng = np.random.default_rng(8)
    for run in fx.runs:
        kw = dict(min_align=run.min_align, pair_flag=int(run.paired), min_dist=run.min_dist, max_dist=run.max_dist,
                  is_bisulfite=int(run.bisulfite))
        n = min(20000, run.reads1.shape[0])
        r1 = run.reads1[:n].copy()
        r2 = run.reads2[:n].copy() if run.paired else None
        L = r1.shape[1]
        for r in (r1, r2):                                  # sprinkle N: a few per read in 2 % of the reads, many in 0.5 %
            if r is not None:
                rows = rng.choice(n, size=n // 50, replace=False)
                for i in rows:
                    r[i, rng.integers(0, L, size=int(rng.integers(1, 4)))] = ord("N")
                for i in rng.choice(n, size=n // 200, replace=False):
                    r[i, rng.integers(0, L, size=L // 8)] = ord("N")
        lens1 = np.full(n, L, dtype=np.int32)
        lens1[::7] = L - 3                                   # ragged: every seventh read loses its last three bases
        lens1[5::11] = max(16, L // 2)
        lens2 = lens1[::-1].copy() if run.paired else None
        out = {}
        for mode in ("rows", "packed"):
            mapper = pb.PEMapper.from_genome(fx.genome, pb.default_params(is_bisulfite=bis))
            mapper.set_params(**kw)
            mapper.keep(pb.KEEP_CANDIDATES)
            if mode == "rows":
                b1, _ = pb.mapper.rows_from_reads(r1)
                b2 = pb.mapper.rows_from_reads(r2)[0] if run.paired else None
                g = mapper.map_rows(b1, lens1, b2, lens2)
            else:
                p1, _ = pb.mapper.pack_reads(r1, lens1, L)
                p2 = pb.mapper.pack_reads(r2, lens2, L)[0] if run.paired else None
                g = mapper.map_packed(p1, lens1, p2, lens2, L)
            cands = [mapper.candidates(i, m) for i in range(0, n, 5) for m in range(2 if run.paired else 1)]
            rec, ins = mapper.finish()
            out[mode] = (g, cands, rec.tobytes(), sorted(ins))
            mapper.close()
        tag = "%s/%s" % (name, run.name)
        for x, y in zip(out["rows"][0], out["packed"][0]):
            assert np.array_equal(x, y), tag + ": per-read results"
        for (s0, o0), (s1, o1) in zip(out["rows"][1], out["packed"][1]):
            assert np.array_equal(s0, s1) and np.array_equal(o0, o1), tag + ": candidate lists"
        assert out["rows"][2] == out["packed"][2], tag + ": pileup records"
        assert out["rows"][3] == out["packed"][3], tag + ": insertions"
        assert (out["rows"][0][0] != 0).sum() > 0.3 * n


def test_cfg4_long_windows_match_oracle(oracle_built):
    """BASELINE configs[3]: reads of 100 / 150 / 250 bp against 1000-bp windows through pemap_sw_score_device.  The
    reference's 300 x 300 buffers cannot hold that shape, so the checker is the oracle's restatement of
    smith_waterman_align with enlarged buffers (orc_sw_align_long, equal to the pinned orc_sw_align wherever both
    apply): score bit-exact in units of 1/36 for every pair, start row / state equal unless the kernel flags an exact
    tie of the last-column maximum.  Shapes include ragged windows (1 .. 1056 rows) and windows at the genome's end."""
    import torch
    from pecaller_b200 import synth
    dev = torch.device("cuda", 0)
    genome = synth.random_genome(44, [300_000])
    g = genome[0]
    mapper = pb.PEMapper.from_genome(genome, pb.default_params(pair_flag=0))
    oracle = ol.Oracle(genome)
    rng = np.random.default_rng(4)
    for L in (100, 150, 250):
        n = 600
        win_len = rng.integers(1, 1057, size=n).astype(np.int32)
        win_len[:200] = 1000
        win_len[200:230] = L + 21
        ws = rng.integers(0, g.shape[0] - 1100, size=n).astype(np.int32)
        ws[-1] = g.shape[0] - win_len[-1]
        stride = (L + 15) // 16 * 16
        reads = np.zeros((n, stride), dtype=np.uint8)
        for i in range(n):
            w = g[ws[i]:ws[i] + win_len[i]]
            if win_len[i] > L + 8 and i % 5:
                o = int(rng.integers(0, win_len[i] - L - 4))
                r = w[o:o + L].copy()
                k = rng.integers(0, L, size=3)
                r[k] = synth.ACGT[rng.integers(0, 4, size=3)]
                if i % 3 == 0:
                    r = np.concatenate([r[:L // 2], r[L // 2 + 2:], w[o + L:o + L + 2]])   # a 2-base deletion in the read
            else:
                r = synth.ACGT[rng.integers(0, 4, size=L)]
            if r.shape[0] < L:
                r = np.concatenate([r, synth.ACGT[rng.integers(0, 4, size=L - r.shape[0])]])
            reads[i, :L] = r[:L]
        t = lambda a: torch.from_numpy(a).to(dev)
        d_reads, d_len, d_ws, d_wl = t(reads), t(np.full(n, L, np.int32)), t(ws), t(win_len)
        sc, mi, mk, fl = (torch.zeros(n, dtype=torch.int32, device=dev) for _ in range(4))
        torch.cuda.synchronize()
        mapper.sw_score_device(n, d_reads.data_ptr(), d_len.data_ptr(), stride, L, d_ws.data_ptr(), d_wl.data_ptr(), 1056,
                               sc.data_ptr(), mi.data_ptr(), mk.data_ptr(), fl.data_ptr())
        sc, mi, mk, fl = (x.cpu().numpy() for x in (sc, mi, mk, fl))
        for i in range(n):
            s, (k, ii, _) = oracle.sw_align_long(int(ws[i]), int(win_len[i]), reads[i, :L].tobytes())
            assert round(s * 36) == sc[i], "L=%d pair %d (window %d): score %r vs %d/36" % (L, i, win_len[i], s, sc[i])
            if not fl[i] & 1:
                assert (ii, k) == (mi[i], mk[i]), "L=%d pair %d: start (%d,%d) vs (%d,%d)" % (L, i, ii, k, mi[i], mk[i])
    mapper.close()
    oracle.close()


def test_contig_count_quirk_is_refused():
    """2..7 contigs: find_chrom reads out of bounds in the reference; we refuse instead of guessing."""
    from pecaller_b200 import synth
    g = synth.random_genome(3, [5000, 5000, 5000])
    with pytest.raises(pb.PemapError):
        pb.PEMapper.from_genome(g)
