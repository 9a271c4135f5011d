"""Host-side multi-GPU plumbing of the hot path: one process per GPU, reads sharded, counters summed once.

SURVEY.md section 8e: reads (pairs) are independent units and the only shared mutable state is the pileup counter
array, a commutative integer sum.  Batch b of `batch` reads goes to rank b % world; every rank maps its batches
against its own replica of the index into a private `uint32 counts[genome_size][6]`; before the writer rank calls
pemap_finish the shards are summed onto it with ONE reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).
Counters are uint32 on the device and truncated to the reference's unsigned short once, after the sum
((sum mod 2^32) mod 2^16 == sum mod 2^16), so 1, 2, 4 and 8 GPUs give byte-identical pileup records.
`m1/m2/mapping_type` stay with the rank that mapped the read; gather_results() rebuilds the per-file arrays
(.mfile order) on the writer rank.  Nothing here touches the oracle or any CPU implementation of the path.
"""
from __future__ import annotations

import numpy as np


def shard_batches(n_reads: int, rank: int, world: int, batch: int):
    """[(start, stop)] of the batches this rank maps: batch b -> rank b % world (round robin)."""
    if batch <= 0 or world <= 0 or not 0 <= rank < world:
        raise ValueError("bad shard arguments")
    out = []
    for b, start in enumerate(range(0, n_reads, batch)):
        if b % world == rank:
            out.append((start, min(n_reads, start + batch)))
    return out


def reduce_counts(counts, dst: int = 0, group=None):
    """Sum the per-rank counter arrays onto rank `dst` in place.  `counts` is an int32 torch tensor viewing the
    uint32 counters (two's-complement addition is the same mod-2^32 sum)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(counts, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return counts


def contig_word_bounds(contig_lens):
    """[(first_word, end_word)] of every contig's slice of the `uint32 counts[genome_size][6]` array (real coordinates:
    contigs are concatenated without separators, pemapper.c:453-494)."""
    out, at = [], 0
    for L in contig_lens:
        out.append((6 * at, 6 * (at + int(L))))
        at += int(L)
    return out


def reduce_counts_by_contig(counts, bounds, dst: int = 0, group=None):
    """The same sum, one chromosome at a time ("per-chromosome pileup arrays are summed with an NCCL
    reduce"): 24 reduces of 1-6 GB each on a human-sized genome instead of one of 74 GB, so that the writer can start
    on chr1 while the later chromosomes are still in flight and NCCL's staging stays small.  Returns the work handles
    (async_op) in chromosome order; wait on handle i before compacting chromosome i."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return []
    return [dist.reduce(counts[a:b], dst=dst, op=dist.ReduceOp.SUM, group=group, async_op=True) for a, b in bounds if b > a]


class CountReducer:
    """Per-rank helper: torch view of the library's counter array + the chromosome slices."""

    def __init__(self, mapper, device, contigs):
        self.counts = counts_tensor(mapper, device)
        self.bounds = contig_word_bounds([c.shape[0] for c in contigs])

    def reduce(self, dst: int = 0):
        for w in reduce_counts_by_contig(self.counts, self.bounds, dst=dst):
            w.wait()


class SliceReducer:
    """One process per GPU: the counters are summed slice-wise over NVLink peer memory (pemap_reduce_scatter_ipc) - every
    rank pulls the other ranks' counters of ITS 1/world of the genome and then compacts that slice itself, so neither
    the sum nor the writer's compaction is serialised on one GPU.  The 64-byte IPC handles are exchanged once."""

    def __init__(self, mapper, group=None):
        import torch.distributed as dist
        self.mapper = mapper
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        handles = [None] * self.world
        dist.all_gather_object(handles, mapper.counts_ipc_handle(), group=group)
        self.handles = handles

    def reduce_scatter(self):
        """-> (site_first, site_end) of this rank's slice, now holding the global sum.  Collective: every rank calls it."""
        import torch.distributed as dist
        dist.barrier(group=self.group)          # every rank has finished mapping into its own array
        rng = self.mapper.reduce_scatter_ipc(self.handles, self.rank)
        dist.barrier(group=self.group)          # nobody resets or maps again while a peer is still reading
        return rng


def gather_results(n_reads: int, ranges, m1, m2, mapping_type, dst: int = 0, group=None):
    """Rebuild the per-read arrays of the whole input on rank `dst` from every rank's shard results.
    ranges: this rank's [(start, stop)] from shard_batches; m1/m2/mapping_type: its results, concatenated in that
    order.  Returns (m1, m2, mapping_type) int64 numpy arrays on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    full = torch.zeros((3, n_reads), dtype=torch.int64)
    at = 0
    for a, b in ranges:
        k = b - a
        full[0, a:b] = torch.as_tensor(np.asarray(m1[at:at + k], dtype=np.int64))
        full[1, a:b] = torch.as_tensor(np.asarray(m2[at:at + k], dtype=np.int64))
        full[2, a:b] = torch.as_tensor(np.asarray(mapping_type[at:at + k], dtype=np.int64))
        at += k
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        t = full.cuda() if backend == "nccl" else full
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)  # shards are disjoint: the sum is a scatter
        full = t.cpu()
        if dist.get_rank(group) != dst:
            return None
    return full[0].numpy(), full[1].numpy(), full[2].numpy()


def counts_tensor(mapper, device):
    """torch int32 view (no copy) of a PEMapper's device counter array, for reduce_counts()."""
    import torch
    ptr, words = mapper.counts_device()

    class _Alias:
        __cuda_array_interface__ = {"shape": (words,), "typestr": "<i4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Alias(), device=device)


def records_from_counts(counts: np.ndarray) -> np.ndarray:
    """Host restatement of the compaction in pemap_finish (writer loop pemapper.c:828-842) for a dense
    [genome_size, 6] counter array: used by the CPU tests of the reduce; the product compacts on the GPU."""
    c16 = (np.asarray(counts).astype(np.int64) & 0xFFFF).astype(np.uint16).reshape(-1, 6)
    pos = np.nonzero(c16.astype(np.uint32).sum(axis=1) > 0)[0]
    rec = np.zeros(pos.shape[0], dtype=np.dtype([("pos", "<u4"), ("c", "<u2", (6,))]))
    rec["pos"] = pos
    rec["c"] = c16[pos]
    return rec
