"""GPU + big host: BASELINE.md section 3.4.  A prefix of the north-star configuration (3.1 Gb genome, 24 contigs) is
mapped by the CUDA path and by the unmodified reference (libpemapper_ref.so) and must agree bit for bit: every
read's .mfile values and mapping type, the sha256 of the pileup records, the insertion multiset.  The committed record
of the full 2 M-pair run is profiles/cfg3_parity_r02.json (tools/cfg3_parity.py); this test repeats it on a shorter
prefix so that the GPU suite stays within minutes.  Skipped without ~170 GB of host RAM or the reference library."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_ram_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def test_cfg3_prefix_equals_reference():
    import oracle_lib as ol
    if not ol.have_reference_lib():
        pytest.skip("oracle/_ref/libpemapper_ref.so was not built")
    if _host_ram_gb() < 170:
        pytest.skip("the reference needs 32 B per genome base (99 GB) + the 16 GiB table on the host")
    if os.environ.get("PEMAP_SKIP_CFG3") == "1":
        pytest.skip("PEMAP_SKIP_CFG3=1")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import cfg3_parity
    out = cfg3_parity.run(int(os.environ.get("PEMAP_CFG3_TEST_PAIRS", 500_000)))
    assert out["identical"]["m1"] and out["identical"]["m2"], "a read maps elsewhere than in the reference"
    assert out["identical"]["mapping_type"]
    assert out["identical"]["pileup_records"], "pileup records differ from the reference's"
    assert out["identical"]["insertions"]
    assert out["cuda"]["type_counts"][0] > 0.99 * out["pairs"]
