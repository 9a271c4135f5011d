"""CPU (gloo, world_size 2): the host-side multi-GPU logic of pecaller_b200.sharding - round-robin batch shards,
the one counter reduce, the result gather - reproduces the single-process result byte for byte.  The per-rank mapper
is the oracle here (test infrastructure); on GPUs it is the CUDA library (tests/test_gpu_parity.py::test_two_gpus)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fixtures_def
import oracle_lib as ol
from pecaller_b200 import sharding


def test_shard_batches_partition():
    for n, world, batch in [(0, 2, 10), (1, 2, 10), (95, 2, 10), (100, 3, 7), (20000, 8, 333)]:
        seen = np.zeros(n, dtype=np.int32)
        for r in range(world):
            for a, b in sharding.shard_batches(n, r, world, batch):
                assert 0 <= a < b <= n and a % batch == 0
                seen[a:b] += 1
        assert (seen == 1).all()
    with pytest.raises(ValueError):
        sharding.shard_batches(10, 2, 2, 5)


def test_records_from_counts_truncates_once():
    c = np.zeros((5, 6), dtype=np.uint32)
    c[1, 0] = 65536 + 3          # wraps to 3 like the reference's unsigned short (pemapper.c:53-58)
    c[2, 4] = 65536              # wraps to 0 -> site not covered
    c[4, 5] = 7
    rec = sharding.records_from_counts(c.view(np.int32))
    assert rec["pos"].tolist() == [1, 4]
    assert rec["c"][0].tolist() == [3, 0, 0, 0, 0, 0] and rec["c"][1][5] == 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dense_counts(oracle):
    rec = oracle.records()
    dense = np.zeros((oracle.genome_size, 6), dtype=np.int32)
    dense[rec["pos"]] = rec["c"]
    return dense


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = fixtures_def.FIXTURES["tiny"]()
    run = fx.runs[1]  # paired
    n = run.reads1.shape[0]
    oracle = ol.Oracle(fx.genome, ol.default_params(min_align=run.min_align, pair_flag=1, min_dist=run.min_dist,
                                                    max_dist=run.max_dist))
    ranges = sharding.shard_batches(n, rank, world, 137)
    parts = [oracle.map_batch(run.reads1[a:b], run.reads2[a:b]) for a, b in ranges]
    m1 = np.concatenate([p[0] for p in parts])
    m2 = np.concatenate([p[1] for p in parts])
    ty = np.concatenate([p[2] for p in parts])
    counts = torch.from_numpy(_dense_counts(oracle).reshape(-1))
    counts += (1 << 16) * (rank + 1)      # every rank's array is shifted by a multiple of 2^16: must vanish
    # chromosome by chromosome, as bench.py and the 8-GPU runs do it (the tiny genome is one contig: cut it in three)
    gs = counts.shape[0] // 6
    bounds = sharding.contig_word_bounds([gs // 3, gs // 3, gs - 2 * (gs // 3)])
    assert bounds[0][0] == 0 and bounds[-1][1] == counts.shape[0]
    for w in sharding.reduce_counts_by_contig(counts, bounds, dst=0):
        w.wait()
    res = sharding.gather_results(n, ranges, m1, m2, ty, dst=0)
    if rank == 0:
        rec = sharding.records_from_counts(counts.numpy())
        np.savez(out_path, rec=rec, m1=res[0], m2=res[1], ty=res[2], ins=np.array(sorted(oracle.insertions()), dtype=object))
    else:
        assert res is None
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path, oracle_built):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out, allow_pickle=True)
    fx = fixtures_def.FIXTURES["tiny"]()
    run = fx.runs[1]
    oracle = ol.Oracle(fx.genome, ol.default_params(min_align=run.min_align, pair_flag=1, min_dist=run.min_dist,
                                                    max_dist=run.max_dist))
    m1, m2, ty = oracle.map_batch(run.reads1, run.reads2)
    assert np.array_equal(got["m1"], m1) and np.array_equal(got["m2"], m2) and np.array_equal(got["ty"], ty)
    assert got["rec"].tobytes() == oracle.records().tobytes()
