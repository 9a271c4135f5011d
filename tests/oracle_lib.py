"""ctypes bindings for the CHECKERS (test infrastructure, never imported by the product):
  * oracle/liboracle.so           - our CPU restatement (oracle/pemap_oracle.c)
  * oracle/_ref/libpemapper_ref.so - the unmodified reference built in-process (oracle/ref_wrap.c), optional
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

TYPE_NAMES = ["UNIQUE_MATE", "UNIQUE_SLIP", "UNIQUE_SINGLE", "UNIQUE_MIS", "NON_MATE", "NON_MIS", "FRAG_MIS", "NON_NO",
              "NEITHER_MAP"]

REC_DTYPE = np.dtype([("pos", "<u4"), ("c", "<u2", (6,))])


class Params(C.Structure):
    _fields_ = [("idepth", C.c_int), ("max_hits", C.c_int), ("too_many_spots", C.c_int), ("min_align", C.c_double),
                ("match_bonus", C.c_double), ("is_bisulfite", C.c_int), ("pair_flag", C.c_int), ("min_dist", C.c_int),
                ("max_dist", C.c_int), ("misalign_slop", C.c_int)]


class Detail(C.Structure):
    _fields_ = [("hits1", C.c_int), ("hits2", C.c_int), ("best1", C.c_int), ("best2", C.c_int), ("orient1", C.c_int),
                ("orient2", C.c_int), ("score1", C.c_double), ("score2", C.c_double)]


DETAIL_DTYPE = np.dtype([("hits1", "<i4"), ("hits2", "<i4"), ("best1", "<i4"), ("best2", "<i4"), ("orient1", "<i4"),
                         ("orient2", "<i4"), ("score1", "<f8"), ("score2", "<f8")])


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_char_p, C.POINTER(C.c_int64), C.c_int, C.POINTER(Params)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.orc_n_mers.restype = C.c_uint64
        L.orc_n_mers.argtypes = [C.c_void_p]
        L.orc_mers.restype = C.POINTER(C.c_uint32)
        L.orc_mers.argtypes = [C.c_void_p]
        L.orc_pos_index.restype = C.c_uint32
        L.orc_pos_index.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_genome_size.restype = C.c_int64
        L.orc_genome_size.argtypes = [C.c_void_p]
        L.orc_contig_starts.restype = C.POINTER(C.c_uint32)
        L.orc_contig_starts.argtypes = [C.c_void_p]
        L.orc_fill_pos_index.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_write_index.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
        L.orc_initial_map.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_window.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
        L.orc_sw_align.restype = C.c_double
        L.orc_sw_align.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_sw_align_long.restype = C.c_double
        L.orc_sw_align_long.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_map_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_count_sites.restype = C.c_uint64
        L.orc_count_sites.argtypes = [C.c_void_p]
        L.orc_get_records.restype = C.c_uint64
        L.orc_get_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_n_insertions.restype = C.c_uint64
        L.orc_n_insertions.argtypes = [C.c_void_p]
        L.orc_get_insertion.restype = C.c_uint32
        L.orc_get_insertion.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_int]
        L.orc_reset_counts.argtypes = [C.c_void_p]
        L.orc_write_indel_txt.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]
        L.orc_cells.restype = C.c_uint64
        L.orc_cells.argtypes = [C.c_void_p]
        L.orc_default_params.argtypes = [C.POINTER(Params)]
        _lib = L
    return _lib


def default_params(**kw) -> Params:
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pad_reads(reads: np.ndarray, stride: int | None = None):
    """(n, L) uint8 -> contiguous (n, stride) NUL-padded matrix + int32 lengths."""
    n, L = reads.shape
    stride = stride or (L + 1)
    buf = np.zeros((n, stride), dtype=np.uint8)
    buf[:, :L] = reads
    return buf, np.full(n, L, dtype=np.int32)


class Oracle:
    def __init__(self, genome, params: Params | None = None):
        self.L = lib()
        self.params = params or default_params()
        self.genome = genome
        cat = np.concatenate(genome)
        lens = (C.c_int64 * len(genome))(*[int(g.shape[0]) for g in genome])
        self.ctx = self.L.orc_create(cat.tobytes(), lens, len(genome), C.byref(self.params))
        self.n_contigs = len(genome)

    def close(self):
        if self.ctx:
            self.L.orc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        self.L.orc_set_params(self.ctx, C.byref(self.params))

    @property
    def genome_size(self):
        return self.L.orc_genome_size(self.ctx)

    def mers(self) -> np.ndarray:
        n = self.L.orc_n_mers(self.ctx)
        return np.ctypeslib.as_array(self.L.orc_mers(self.ctx), shape=(n,)).copy()

    def contig_starts(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.orc_contig_starts(self.ctx), shape=(self.n_contigs + 1,)).copy()

    def pos_index(self, w: int) -> int:
        return self.L.orc_pos_index(self.ctx, w)

    def dense_pos_index(self) -> np.ndarray:
        """The inflated .idx: 2^32+1 uint32 (16 GiB)."""
        t = np.empty((1 << 32) + 1, dtype=np.uint32)
        self.L.orc_fill_pos_index(self.ctx, t.ctypes.data)
        return t

    def write_index(self, base: str, names, with_idx=False):
        arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
        return self.L.orc_write_index(self.ctx, base.encode(), arr, int(with_idx))

    def initial_map(self, read: bytes):
        spots = np.zeros(256, dtype=np.uint32)
        orients = np.zeros(256, dtype=np.int8)
        n = self.L.orc_initial_map(self.ctx, read, len(read), spots.ctypes.data, orients.ctypes.data)
        return spots[:n].copy(), orients[:n].copy()

    def window(self, spot: int, length: int):
        s = C.c_uint32()
        b = C.c_int()
        ch = self.L.orc_window(self.ctx, spot, length, C.byref(s), C.byref(b))
        return ch, s.value, b.value

    def sw_align(self, win_start: int, blen: int, seq: bytes):
        st = (C.c_int * 3)()
        sc = self.L.orc_sw_align(self.ctx, win_start, blen, seq, len(seq), st)
        return sc, (st[0], st[1], st[2])

    def sw_align_long(self, win_start: int, blen: int, seq: bytes):
        """the same recurrence for windows beyond the reference's 300 x 300 buffers (cfg4's 1000-bp windows)"""
        st = (C.c_int * 3)()
        sc = self.L.orc_sw_align_long(self.ctx, win_start, blen, seq, len(seq), st)
        return sc, (st[0], st[1], st[2])

    def map_batch(self, reads1, reads2=None, nthreads=1, detail=False):
        b1, l1 = pad_reads(reads1, 304)
        n = b1.shape[0]
        b2 = l2 = None
        if reads2 is not None:
            b2, l2 = pad_reads(reads2, 304)
        m1 = np.zeros(n, dtype=np.uint32)
        m2 = np.zeros(n, dtype=np.uint32)
        ty = np.zeros(n, dtype=np.int32)
        det = np.zeros(n, dtype=DETAIL_DTYPE) if detail else None
        self.L.orc_map_batch(self.ctx, n, b1.ctypes.data, l1.ctypes.data, b2.ctypes.data if b2 is not None else None,
                             l2.ctypes.data if l2 is not None else None, 304, m1.ctypes.data, m2.ctypes.data,
                             ty.ctypes.data, det.ctypes.data if detail else None, nthreads)
        return (m1, m2, ty, det) if detail else (m1, m2, ty)

    def records(self) -> np.ndarray:
        n = self.L.orc_count_sites(self.ctx)
        out = np.zeros(n, dtype=REC_DTYPE)
        self.L.orc_get_records(self.ctx, out.ctypes.data, n)
        return out

    def insertions(self):
        n = self.L.orc_n_insertions(self.ctx)
        buf = C.create_string_buffer(512)
        out = []
        for i in range(n):
            pos = self.L.orc_get_insertion(self.ctx, i, buf, 512)
            out.append((pos, buf.value.decode()))
        return sorted(out)

    def indel_txt(self, names, path):
        arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
        self.L.orc_write_indel_txt(self.ctx, path.encode(), arr)
        return open(path).read()

    def reset(self):
        self.L.orc_reset_counts(self.ctx)

    def cells(self):
        return self.L.orc_cells(self.ctx)


# ---------------------------------------------------------------- the unmodified reference, in-process

REF_SO = os.path.join(ORACLE_DIR, "_ref", "libpemapper_ref.so")


def have_reference_lib():
    return os.path.exists(REF_SO)


class ReferenceLib:
    """One per process (the reference keeps its state in file-static globals)."""
    _inst = None

    def __init__(self, oracle: Oracle | None, min_align=0.9, bisulfite=False, paired=False, min_dist=0, max_dist=500,
                 arrays=None):
        """arrays = (genome_cat, contig_starts, pos_index[2^32+1], mers, n_contigs): an index built elsewhere (the
        device index builder, bit-equal to index_genome_whole's files) instead of the oracle's CPU indexer - the
        only practical way to a 3.1 Gb genome."""
        assert ReferenceLib._inst is None, "the reference library can be initialised once per process"
        ReferenceLib._inst = self
        R = C.CDLL(REF_SO)
        self.R = R
        R.refw_init.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                C.c_int, C.c_int, C.c_int]
        R.refw_set_params.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int]
        R.refw_map.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int]
        R.refw_records.restype = C.c_ulong
        R.refw_records.argtypes = [C.c_void_p, C.c_ulong]
        R.refw_n_insertions.restype = C.c_ulong
        R.refw_insertions.restype = C.c_ulong
        R.refw_insertions.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_ulong]
        R.refw_initial_map.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        R.refw_sw_align.restype = C.c_double
        R.refw_sw_align.argtypes = [C.c_uint, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        for f in ("refw_mate_count", "refw_total_reads", "refw_total_bases", "refw_total_dist", "refw_no_dists"):
            getattr(R, f).restype = C.c_long
        R.refw_last_live_threads.restype = C.c_int
        if arrays is not None:
            self.genome, self.cstarts, self.pos_index, self.mers, n_contigs = arrays
            assert self.pos_index.dtype == np.uint32 and self.pos_index.shape[0] == (1 << 32) + 1
            assert self.mers.dtype == np.uint32 and self.cstarts.dtype == np.uint32
        else:
            self.genome = np.concatenate(oracle.genome)
            self.cstarts = oracle.contig_starts()
            self.pos_index = oracle.dense_pos_index()
            self.mers = np.concatenate([oracle.mers(), np.zeros(4, np.uint32)])
            n_contigs = oracle.n_contigs
        rc = R.refw_init(self.genome.ctypes.data, self.genome.shape[0], self.cstarts.ctypes.data, n_contigs,
                         self.pos_index.ctypes.data, self.mers.ctypes.data, min_align, int(bisulfite), int(paired),
                         min_dist, max_dist)
        assert rc == 0

    def set_params(self, min_align, paired, min_dist=0, max_dist=500):
        self.R.refw_set_params(min_align, int(paired), min_dist, max_dist)

    def map(self, reads1, reads2=None, nthreads=1):
        b1, l1 = pad_reads(reads1, 304)
        n = b1.shape[0]
        b2 = l2 = None
        if reads2 is not None:
            b2, l2 = pad_reads(reads2, 304)
        m1 = np.zeros(n, dtype=np.uint32)
        m2 = np.zeros(n, dtype=np.uint32)
        ty = np.zeros(n, dtype=np.int32)
        rc = self.R.refw_map(n, b1.ctypes.data, l1.ctypes.data, b2.ctypes.data if b2 is not None else None,
                             l2.ctypes.data if l2 is not None else None, 304, m1.ctypes.data, m2.ctypes.data,
                             ty.ctypes.data, nthreads)
        assert rc == 0
        return m1, m2, ty

    def live_threads(self):
        """worker threads that had a batch in the widest wave of the last map() call"""
        return self.R.refw_last_live_threads()

    def records(self):
        n = self.R.refw_records(None, 0)
        out = np.zeros(n, dtype=REC_DTYPE)
        self.R.refw_records(out.ctypes.data, n)
        return out

    def insertions(self):
        n = self.R.refw_n_insertions()
        pos = np.zeros(n, dtype=np.uint32)
        buf = np.zeros((n, 304), dtype=np.uint8)
        self.R.refw_insertions(pos.ctypes.data, buf.ctypes.data, 304, n)
        return sorted((int(p), bytes(b).split(b"\0")[0].decode()) for p, b in zip(pos, buf))

    def reset(self):
        self.R.refw_reset_counts()

    def initial_map(self, read: bytes):
        spots = np.zeros(256, dtype=np.uint32)
        orients = np.zeros(256, dtype=np.int8)
        n = self.R.refw_initial_map(read, len(read), spots.ctypes.data, orients.ctypes.data)
        return spots[:n].copy(), orients[:n].copy()

    def sw_align(self, win_start, blen, seq: bytes):
        st = (C.c_int * 3)()
        sc = self.R.refw_sw_align(win_start, blen, seq, len(seq), st)
        return sc, (st[0], st[1], st[2])

    def mate_counts(self):
        return [self.R.refw_mate_count(i) for i in range(9)]

    def totals(self):
        return dict(total_reads=self.R.refw_total_reads(), total_bases=self.R.refw_total_bases(),
                    total_dist=self.R.refw_total_dist(), no_dists=self.R.refw_no_dists())
