"""ctypes binding of the C-ABI (include/pemap.h) and a thin host-side mirror of the reference's batch worker.

The reference has no Python; this mirror exists so that tests and bench.py can drive the C-ABI exactly the
way a patched pemapper.c would: init once, map_batch per block of reads, finish once.
There is no CPU fallback anywhere: if the CUDA library is missing or no B200 is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

TYPE_NAMES = ["UNIQUE_MATE", "UNIQUE_SLIP", "UNIQUE_SINGLE", "UNIQUE_MIS", "NON_MATE", "NON_MIS", "FRAG_MIS", "NON_NO",
              "NEITHER_MAP"]
RECORD_DTYPE = np.dtype([("pos", "<u4"), ("c", "<u2", (6,))])
DETAIL_DTYPE = np.dtype([("hits1", "<i4"), ("hits2", "<i4"), ("best1", "<i4"), ("best2", "<i4"), ("orient1", "<i4"),
                         ("orient2", "<i4"), ("score1", "<f8"), ("score2", "<f8")])
KEEP_DETAIL = 1
KEEP_CANDIDATES = 2


class PemapError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("idepth", C.c_int), ("max_hits", C.c_int), ("too_many_spots", C.c_int), ("min_align", C.c_double),
                ("match_bonus", C.c_double), ("is_bisulfite", C.c_int), ("pair_flag", C.c_int), ("min_dist", C.c_int),
                ("max_dist", C.c_int), ("misalign_slop", C.c_int)]


class Index(C.Structure):
    _fields_ = [("pos_index", C.c_void_p), ("mers", C.c_void_p), ("n_mers", C.c_uint64), ("genome", C.c_void_p),
                ("genome_size", C.c_uint64), ("contig_starts", C.c_void_p), ("no_contigs", C.c_int)]


class Insertion(C.Structure):
    _fields_ = [("pos", C.c_uint32), ("len", C.c_uint32), ("seq", C.c_char_p)]


class Stats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("lookups", C.c_uint64), ("mer_positions", C.c_uint64),
                ("candidates", C.c_uint64), ("sw_cells", C.c_uint64), ("tb_cells", C.c_uint64),
                ("replayed", C.c_uint64), ("ms_seed", C.c_double), ("ms_sw", C.c_double), ("ms_select", C.c_double),
                ("ms_traceback", C.c_double), ("ms_total", C.c_double), ("launches", C.c_uint64),
                ("diag_traced", C.c_uint64), ("exact_traced", C.c_uint64),
                ("ms_tb_diag", C.c_double), ("ms_tb_int", C.c_double), ("ms_tb_fp64", C.c_double),
                ("tb_cells_int", C.c_uint64), ("sw_cells_certified", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64)  # pemap_site_cb(ctx, records, n)

EXPORTS = ["pemap_version", "pemap_default_params", "pemap_init", "pemap_init_streamed", "pemap_init_from_genome", "pemap_set_params",
           "pemap_map_batch", "pemap_map_batch_rows", "pemap_map_batch_device", "pemap_keep", "pemap_get_detail",
           "pemap_get_candidates", "pemap_finish", "pemap_finish_stream", "pemap_finish_stream_range", "pemap_get_insertions", "pemap_counts_ipc_handle",
           "pemap_reduce_scatter_ipc", "pemap_reduce_scatter_local", "pemap_host_alloc", "pemap_host_free", "pemap_sw_score_device", "pemap_packed_stride", "pemap_pack_read",
           "pemap_map_batch_packed", "pemap_reset_counts", "pemap_counts_device", "pemap_get_stats",
           "pemap_reset_stats", "pemap_stream", "pemap_reduce_counts_peer", "pemap_index_device", "pemap_read_pos_index", "pemap_read_mers", "pemap_last_error",
           "pemap_destroy"]

_lib = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """Load pecaller_b200/libpemap.so (building it with nvcc when absent)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PEMAP_LIB") or _build.LIB  # PEMAP_LIB: an instrumented build (e.g. -DPM_TIE_DEBUG)
    if not os.path.exists(path):
        if not build_if_missing:
            raise PemapError("CUDA library %s is missing; run __graft_entry__.build()" % path)
        _build.build()
    L = C.CDLL(path)
    vp, u32p, ip = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int)
    L.pemap_version.restype = C.c_char_p
    L.pemap_last_error.restype = C.c_char_p
    L.pemap_last_error.argtypes = [vp]
    L.pemap_default_params.argtypes = [C.POINTER(Params)]
    L.pemap_init.argtypes = [C.POINTER(vp), C.POINTER(Index), C.POINTER(Params), C.c_int]
    L.pemap_init_from_genome.argtypes = [C.POINTER(vp), vp, C.POINTER(C.c_int64), C.c_int, C.POINTER(Params), C.c_int]
    L.pemap_set_params.argtypes = [vp, C.POINTER(Params)]
    L.pemap_map_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.pemap_map_batch_rows.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp]
    L.pemap_map_batch_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.pemap_keep.argtypes = [vp, C.c_int]
    L.pemap_get_detail.argtypes = [vp, vp, C.c_int]
    L.pemap_get_candidates.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int]
    L.pemap_finish.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.POINTER(Insertion)),
                               C.POINTER(C.c_uint64)]
    L.pemap_finish_stream.argtypes = [vp, SITE_CB, vp, C.POINTER(C.c_uint64)]
    L.pemap_finish_stream_range.argtypes = [vp, C.c_uint64, C.c_uint64, SITE_CB, vp, C.POINTER(C.c_uint64)]
    L.pemap_counts_ipc_handle.argtypes = [vp, vp]
    L.pemap_reduce_scatter_ipc.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pemap_reduce_scatter_local.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.pemap_sw_score_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp, C.POINTER(C.c_float)]
    L.pemap_packed_stride.restype = C.c_size_t
    L.pemap_packed_stride.argtypes = [C.c_int]
    L.pemap_pack_read.argtypes = [vp, C.c_int, C.c_int, vp]
    L.pemap_map_batch_packed.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp]
    L.pemap_host_alloc.restype = vp
    L.pemap_host_alloc.argtypes = [C.c_size_t]
    L.pemap_host_free.argtypes = [vp]
    L.pemap_get_insertions.argtypes = [vp, C.POINTER(C.POINTER(Insertion)), C.POINTER(C.c_uint64)]
    L.pemap_reset_counts.argtypes = [vp]
    L.pemap_counts_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.pemap_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.pemap_reset_stats.argtypes = [vp]
    L.pemap_stream.argtypes = [vp, C.POINTER(vp)]
    L.pemap_reduce_counts_peer.argtypes = [vp, vp]
    L.pemap_index_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.pemap_read_pos_index.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.pemap_read_mers.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.pemap_destroy.argtypes = [vp]
    _lib = L
    return L


def default_params(**kw) -> Params:
    p = Params()
    load_library().pemap_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def rows_from_reads(reads: np.ndarray, stride: int | None = None):
    """(n, L) uint8 reads -> ((n, stride) row matrix, int32 lengths) as the C-ABI's *_rows entry takes them."""
    n, L = reads.shape
    stride = stride or ((L + 15) // 16 * 16)
    if stride == L:
        return np.ascontiguousarray(reads), np.full(n, L, dtype=np.int32)
    buf = np.zeros((n, stride), dtype=np.uint8)
    buf[:, :L] = reads
    return buf, np.full(n, L, dtype=np.int32)


def pack_reads(reads: np.ndarray, lens: np.ndarray | None = None, max_len: int | None = None):
    """(n, L) uint8 ASCII reads -> ((n, pemap_packed_stride(max_len)) uint8 packed rows, int32 lengths), the layout of
    pemap_pack_read, vectorised in numpy (bench and tests; a C caller packs read by read with pemap_pack_read).
    Raises PemapError for characters other than upper-case ACGTN."""
    L = load_library()
    n, width = reads.shape
    lens = np.full(n, width, dtype=np.int32) if lens is None else np.asarray(lens, dtype=np.int32)
    max_len = max_len or int(lens.max() if n else 1)
    stride = L.pemap_packed_stride(max_len)
    code_words, mask_words = (max_len + 15) // 16, (max_len + 31) // 32
    code = np.full(256, 255, dtype=np.uint8)
    for ch, c in zip(b"ACGTN", (0, 1, 2, 3, 0)):
        code[ch] = c
    valid = np.arange(width)[None, :] < lens[:, None]
    c = code[reads]
    if (c[valid] == 255).any():
        raise PemapError("pack_reads: a read holds characters other than ACGTN")
    c = np.where(valid, c, 0).astype(np.uint32)
    isn = (reads == ord("N")) & valid
    pad = (-width) % 32
    if pad:
        c = np.concatenate([c, np.zeros((n, pad), np.uint32)], axis=1)
        isn = np.concatenate([isn, np.zeros((n, pad), bool)], axis=1)
    cw = (c.reshape(n, -1, 16) << (30 - 2 * np.arange(16, dtype=np.uint32))[None, None, :]).sum(axis=2, dtype=np.uint32)
    mw = (isn.reshape(n, -1, 32).astype(np.uint32) << np.arange(32, dtype=np.uint32)[None, None, :]).sum(axis=2, dtype=np.uint32)
    out = np.zeros((n, stride // 4), dtype=np.uint32)
    out[:, :code_words] = cw[:, :code_words]
    out[:, code_words:code_words + mask_words] = mw[:, :mask_words]
    return out.view(np.uint8).reshape(n, stride), lens


class PEMapper:
    """One B200, one replicated index, one private pileup counter array (= one map_everything worker pool)."""

    def __init__(self, handle, lib, params):
        self._h = handle
        self._L = lib
        self.params = params

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_genome(cls, genome, params: Params | None = None, device: int = 0) -> "PEMapper":
        """genome: list of upper-case ASCII contigs (uint8 arrays). Index is built on the device."""
        L = load_library()
        params = params or default_params()
        cat = np.ascontiguousarray(np.concatenate(genome))
        lens = (C.c_int64 * len(genome))(*[int(g.shape[0]) for g in genome])
        h = C.c_void_p()
        rc = L.pemap_init_from_genome(C.byref(h), cat.ctypes.data, lens, len(genome), C.byref(params), device)
        return cls._check_new(L, h, rc, params)

    @classmethod
    def from_index(cls, pos_index: np.ndarray, mers: np.ndarray, genome: np.ndarray, contig_starts: np.ndarray,
                   params: Params | None = None, device: int = 0) -> "PEMapper":
        """The arrays pemapper's main() holds after loading G.idx/.mdx/.seq/.sdx (pemapper.c:411-538)."""
        L = load_library()
        params = params or default_params()
        assert pos_index.dtype == np.uint32 and pos_index.shape[0] == (1 << 32) + 1
        ix = Index(pos_index.ctypes.data, mers.ctypes.data, mers.shape[0], genome.ctypes.data, genome.shape[0],
                   contig_starts.ctypes.data, contig_starts.shape[0] - 1)
        h = C.c_void_p()
        rc = L.pemap_init(C.byref(h), C.byref(ix), C.byref(params), device)
        return cls._check_new(L, h, rc, params)

    @classmethod
    def _check_new(cls, L, h, rc, params):
        if rc != 0:
            msg = L.pemap_last_error(h).decode() if h else "no handle"
            if h:
                L.pemap_destroy(h)
            raise PemapError("pemap_init failed (%d): %s" % (rc, msg))
        return cls(h, L, params)

    def _ck(self, rc):
        if rc < 0:
            raise PemapError("pemap error %d: %s" % (rc, self._L.pemap_last_error(self._h).decode()))
        return rc

    def close(self):
        if self._h:
            self._L.pemap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ mapping
    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        self._ck(self._L.pemap_set_params(self._h, C.byref(self.params)))

    def keep(self, flags: int):
        self._ck(self._L.pemap_keep(self._h, flags))

    def map_batch(self, reads1: np.ndarray, reads2: np.ndarray | None = None):
        """reads: (n, L) uint8 matrices. Returns (m1, m2, mapping_type) like PTHREAD_DATA_NODE.m1/m2/mapping_type."""
        r1, l1 = rows_from_reads(reads1)
        n = r1.shape[0]
        r2 = l2 = None
        if reads2 is not None:
            r2, l2 = rows_from_reads(reads2, r1.shape[1] if reads2.shape[1] == reads1.shape[1] else None)
            if r2.shape[1] != r1.shape[1]:
                s = max(r1.shape[1], r2.shape[1])
                r1, l1 = rows_from_reads(reads1, s)
                r2, l2 = rows_from_reads(reads2, s)
        return self.map_rows(r1, l1, r2, l2)

    def map_rows(self, r1, l1, r2=None, l2=None):
        n = r1.shape[0]
        m1 = np.zeros(n, dtype=np.uint32)
        m2 = np.zeros(n, dtype=np.uint32)
        ty = np.zeros(n, dtype=np.int32)
        self._ck(self._L.pemap_map_batch_rows(self._h, n, r1.ctypes.data, l1.ctypes.data,
                                              r2.ctypes.data if r2 is not None else None,
                                              l2.ctypes.data if l2 is not None else None, r1.shape[1],
                                              m1.ctypes.data, m2.ctypes.data, ty.ctypes.data))
        return m1, m2, ty

    def map_packed(self, p1, l1, p2, l2, max_len):
        """pemap_map_batch_packed: p1/p2 are (n, pemap_packed_stride(max_len)) uint8 matrices from pack_reads()."""
        n = p1.shape[0]
        m1 = np.zeros(n, dtype=np.uint32)
        m2 = np.zeros(n, dtype=np.uint32)
        ty = np.zeros(n, dtype=np.int32)
        self._ck(self._L.pemap_map_batch_packed(self._h, n, p1.ctypes.data, l1.ctypes.data,
                                                p2.ctypes.data if p2 is not None else None,
                                                l2.ctypes.data if l2 is not None else None, max_len, m1.ctypes.data,
                                                m2.ctypes.data, ty.ctypes.data))
        return m1, m2, ty

    def map_pointers(self, reads1: list, reads2: list | None = None):
        """The PTHREAD_DATA_NODE form: arrays of NUL-terminated char* (pemap_map_batch)."""
        n = len(reads1)
        a1 = (C.c_char_p * n)(*reads1)
        l1 = (C.c_int * n)(*[len(r) for r in reads1])
        a2 = l2 = None
        if reads2 is not None:
            a2 = (C.c_char_p * n)(*reads2)
            l2 = (C.c_int * n)(*[len(r) for r in reads2])
        m1 = np.zeros(n, dtype=np.uint32)
        m2 = np.zeros(n, dtype=np.uint32)
        ty = np.zeros(n, dtype=np.int32)
        self._ck(self._L.pemap_map_batch(self._h, n, a1, l1, a2, l2, m1.ctypes.data, m2.ctypes.data, ty.ctypes.data))
        return m1, m2, ty

    def reduce_counts_from(self, other: "PEMapper"):
        """Add another handle's (another GPU's) pileup counters into this one over NVLink peer memory."""
        self._ck(self._L.pemap_reduce_counts_peer(self._h, other._h))

    def map_device(self, n, d_r1, d_l1, d_r2, d_l2, stride, max_len, d_m1, d_m2, d_ty):
        """All arguments are raw device pointers (ints)."""
        self._ck(self._L.pemap_map_batch_device(self._h, n, d_r1, d_l1, d_r2, d_l2, stride, max_len, d_m1, d_m2, d_ty))

    def sw_score_device(self, n, d_reads, d_len, stride, max_len, d_win_start, d_win_len, max_window, d_score36, d_maxi,
                        d_maxk, d_flags=None) -> float:
        """pemap_sw_score_device (raw device pointers); returns the kernel's milliseconds."""
        ms = C.c_float()
        self._ck(self._L.pemap_sw_score_device(self._h, n, d_reads, d_len, stride, max_len, d_win_start, d_win_len, max_window,
                                               d_score36, d_maxi, d_maxk, d_flags, C.byref(ms)))
        return ms.value

    def detail(self, n) -> np.ndarray:
        out = np.zeros(n, dtype=DETAIL_DTYPE)
        self._ck(self._L.pemap_get_detail(self._h, out.ctypes.data, n))
        return out

    def candidates(self, i, mate=0):
        spots = np.zeros(256, dtype=np.uint32)
        orients = np.zeros(256, dtype=np.int8)
        n = self._ck(self._L.pemap_get_candidates(self._h, i, mate, spots.ctypes.data, orients.ctypes.data, 256))
        return spots[:n].copy(), orients[:n].copy()

    # ------------------------------------------------------------------ results
    def finish(self):
        """-> (records: RECORD_DTYPE array in ascending position, insertions: sorted list of (pos, str))."""
        rec = C.c_void_p()
        nrec = C.c_uint64()
        ins = C.POINTER(Insertion)()
        nins = C.c_uint64()
        self._ck(self._L.pemap_finish(self._h, C.byref(rec), C.byref(nrec), C.byref(ins), C.byref(nins)))
        if nrec.value:
            buf = (C.c_char * (16 * nrec.value)).from_address(rec.value)
            records = np.frombuffer(buf, dtype=RECORD_DTYPE).copy()
        else:
            records = np.zeros(0, dtype=RECORD_DTYPE)
        insertions = [(int(ins[i].pos), ins[i].seq.decode()) for i in range(nins.value)]
        return records, insertions

    def finish_stream(self, consume=None, site_range=None):
        """pemap_finish_stream(_range): `consume(records)` is called with every window's RECORD_DTYPE view (valid only
        during the call) in ascending position; returns the number of records.  consume=None only counts."""
        def cb(_ctx, ptr, n):
            if consume is not None:
                buf = (C.c_char * (16 * n)).from_address(ptr)
                consume(np.frombuffer(buf, dtype=RECORD_DTYPE))
            return 0
        total = C.c_uint64()
        if site_range is None:
            self._ck(self._L.pemap_finish_stream(self._h, SITE_CB(cb), None, C.byref(total)))
        else:
            self._ck(self._L.pemap_finish_stream_range(self._h, site_range[0], site_range[1], SITE_CB(cb), None, C.byref(total)))
        return total.value

    def counts_ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self._L.pemap_counts_ipc_handle(self._h, buf))
        return buf.raw

    def reduce_scatter_ipc(self, handles: list, rank: int):
        """handles: every rank's counts_ipc_handle() in rank order.  -> (site_first, site_end) of this rank's slice."""
        blob = b"".join(handles)
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._L.pemap_reduce_scatter_ipc(self._h, blob, len(handles), rank, C.byref(a), C.byref(b)))
        return a.value, b.value

    @staticmethod
    def reduce_scatter_local(mappers: list, which: int):
        L = mappers[0]._L
        arr = (C.c_void_p * len(mappers))(*[m._h for m in mappers])
        a, b = C.c_uint64(), C.c_uint64()
        mappers[which]._ck(L.pemap_reduce_scatter_local(arr, len(mappers), which, C.byref(a), C.byref(b)))
        return a.value, b.value

    def insertions(self):
        ins = C.POINTER(Insertion)()
        nins = C.c_uint64()
        self._ck(self._L.pemap_get_insertions(self._h, C.byref(ins), C.byref(nins)))
        return [(int(ins[i].pos), ins[i].seq.decode()) for i in range(nins.value)]

    def reset_counts(self):
        self._ck(self._L.pemap_reset_counts(self._h))

    def counts_device(self):
        p = C.c_void_p()
        n = C.c_uint64()
        self._ck(self._L.pemap_counts_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stats(self) -> dict:
        s = Stats()
        self._ck(self._L.pemap_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        self._ck(self._L.pemap_stream(self._h, C.byref(p)))
        return p.value or 0

    def reset_stats(self):
        self._ck(self._L.pemap_reset_stats(self._h))

    def read_mers(self) -> np.ndarray:
        n = C.c_uint64()
        self._ck(self._L.pemap_index_device(self._h, None, None, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint32)
        if n.value:
            self._ck(self._L.pemap_read_mers(self._h, 0, n.value, out.ctypes.data))
        return out

    def read_pos_index(self, first, n) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint32)
        self._ck(self._L.pemap_read_pos_index(self._h, first, n, out.ctypes.data))
        return out


def summary_counts(m1, m2, types, len1, len2, max_dist, paired):
    """The batch epilogue of map_everything (pemapper.c:1238-1265) on the host: mate_counts, total_reads,
    total_bases, total_dist, no_dists - including the unsigned |m1-m2| quirk."""
    mate_counts = np.bincount(types, minlength=9).astype(np.int64)
    has1 = m1 != 0
    has2 = m2 != 0
    total_reads = int(has1.sum() + has2.sum())
    total_bases = int(np.where(has1, len1, 0).sum() + (np.where(has2, len2, 0).sum() if paired else 0))
    both = has1 & has2
    test = (m1[both].astype(np.uint32) - m2[both].astype(np.uint32)).astype(np.int64)  # unsigned wrap, then long
    ok = test < 4 * max_dist
    return dict(mate_counts=mate_counts, total_reads=total_reads, total_bases=total_bases,
                total_dist=int(test[ok].sum()), no_dists=int(ok.sum()))
