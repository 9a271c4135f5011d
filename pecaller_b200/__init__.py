"""pecaller_b200 - B200-native (sm_100a) implementation of PEMapper's read-mapping hot path behind a C-ABI.

Package contents: csrc/ (CUDA kernels + C-ABI, built in-tree into libpemap.so), mapper.py (ctypes binding and
host mirror of the reference's batch worker), host/ (C command-line host), synth.py (seeded synthetic data).
"""
from .mapper import (PEMapper, PemapError, Params, default_params, load_library, summary_counts,  # noqa: F401
                     RECORD_DTYPE, DETAIL_DTYPE, TYPE_NAMES, KEEP_DETAIL, KEEP_CANDIDATES, EXPORTS)

__all__ = ["PEMapper", "PemapError", "Params", "default_params", "load_library", "summary_counts", "RECORD_DTYPE",
           "DETAIL_DTYPE", "TYPE_NAMES", "KEEP_DETAIL", "KEEP_CANDIDATES", "EXPORTS"]
