"""CPU: the C-ABI library builds, loads and exports every symbol include/pemap.h declares; without a GPU every
entry point fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import pecaller_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pemap.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pemap_[a-z_]+)\s*\(", txt)))


def test_header_symbols_exported():
    L = pb.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libpemap.so does not export %s" % s
    assert sorted(pb.EXPORTS) == syms


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(pb.Params) == 56
    assert pb.RECORD_DTYPE.itemsize == 16
    assert pb.DETAIL_DTYPE.itemsize == 40


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = [np.frombuffer(b"ACGT" * 50, dtype=np.uint8)]
    with pytest.raises(pb.PemapError):
        pb.PEMapper.from_genome(g)


def test_product_does_not_reference_oracle():
    """The product path must never import, link or load anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pecaller_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "pemap_oracle" not in txt and "oracle_lib" not in txt, f
