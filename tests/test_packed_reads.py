"""CPU: the 2-bit packed read rows of include/pemap.h - pemap_pack_read (what a C caller uses, read by read) against the
vectorised numpy packer the tests and bench.py use, the documented bit layout, and the refusal of characters that the
reference scores by exact character (IUPAC codes, lower case)."""
import ctypes as C

import numpy as np
import pytest

import pecaller_b200 as pb
from pecaller_b200 import mapper as M


def test_pack_read_layout_and_numpy_packer_agree():
    L = pb.load_library()
    rng = np.random.default_rng(5)
    for max_len in (16, 100, 150, 250, 298):
        stride = L.pemap_packed_stride(max_len)
        assert stride % 16 == 0 and stride >= 4 * ((max_len + 15) // 16) + 4 * ((max_len + 31) // 32)
        n = 50
        reads = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.choice(5, size=(n, max_len), p=[0.24, 0.24, 0.24, 0.24, 0.04])]
        lens = rng.integers(0, max_len + 1, size=n).astype(np.int32)
        lens[0], lens[1] = max_len, 0
        packed, _ = M.pack_reads(reads, lens, max_len)
        assert packed.shape == (n, stride)
        cw = (max_len + 15) // 16
        for i in range(n):
            buf = (C.c_uint8 * stride)()
            rc = L.pemap_pack_read(reads[i].tobytes(), int(lens[i]), max_len, buf)
            assert rc == 0
            assert bytes(buf) == packed[i].tobytes(), "row %d of max_len %d" % (i, max_len)
            words = np.frombuffer(bytes(buf), dtype="<u4")
            for j in range(int(lens[i])):            # the documented layout, base by base
                code = (int(words[j // 16]) >> (30 - 2 * (j % 16))) & 3
                isn = (int(words[cw + j // 32]) >> (j % 32)) & 1
                ch = reads[i, j]
                assert isn == (ch == ord("N")) and code == {65: 0, 67: 1, 71: 2, 84: 3, 78: 0}[int(ch)]


def test_pack_refuses_what_the_reference_scores_by_character():
    L = pb.load_library()
    buf = (C.c_uint8 * L.pemap_packed_stride(32))()
    assert L.pemap_pack_read(b"ACGTRACGTACGTACGT", 17, 32, buf) < 0      # IUPAC code
    assert L.pemap_pack_read(b"ACGTaACGTACGTACGT", 17, 32, buf) < 0      # lower case
    assert L.pemap_pack_read(b"ACGT", 40, 32, buf) < 0                    # longer than the row
    with pytest.raises(pb.PemapError):
        M.pack_reads(np.frombuffer(b"ACGTYACGTACGTACGT", dtype=np.uint8)[None, :])
