"""GPU + big host: BASELINE.md section 3.4 and its cfg5 counterpart.  A prefix of the north-star configuration (3.1 Gb
genome, 24 contigs) and reads of the high-repeat genome (cfg5) are mapped by the CUDA path and by the unmodified
reference (oracle/_ref/libpemapper_ref.so, pemapper.c's own map_everything on all host threads) and must agree bit for
bit: every read's .mfile values and mapping type, the sha256 of the pileup records, the insertion multiset.
tools/cfg3_parity.py does the work in a process of its own (the reference keeps its state in file-static globals, so a
process can hold one genome); the committed records of the full-size runs are profiles/cfg3_parity_r02.json and
profiles/cfg5_parity_r02.json.  Skipped without the reference library or enough host RAM."""
import json
import os
import subprocess
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


def _parity(config, n, need_gb, also_single=0):
    import oracle_lib as ol
    if not ol.have_reference_lib():
        pytest.skip("oracle/_ref/libpemapper_ref.so was not built")
    if _host_ram_gb() < need_gb:
        pytest.skip("the reference needs 32 B per genome base + the 16 GiB table on the host")
    if os.environ.get("PEMAP_SKIP_BIG_PARITY") == "1":
        pytest.skip("PEMAP_SKIP_BIG_PARITY=1")
    env = dict(os.environ, PEMAP_PARITY_CONFIG=config, PEMAP_PARITY_SINGLE="0", PEMAP_PARITY_ALSO_SINGLE=str(also_single))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cfg3_parity.py"), str(n)], env=env, capture_output=True,
                       text=True, timeout=1500)
    assert r.stdout.strip(), r.stderr[-2000:]
    out = json.loads(r.stdout)
    for part in (out, out.get("single_end")):
        if part:
            for k, v in part["identical"].items():
                assert v, "%s: %s differs from the reference (%s)" % (config, k, json.dumps(part["cuda"])[:400])
    return out


def test_cfg3_prefix_equals_reference():
    out = _parity("cfg3", int(os.environ.get("PEMAP_CFG3_TEST_PAIRS", 300_000)), 170)
    assert out["cuda"]["type_counts"][0] > 0.99 * out["pairs"]


def test_cfg5_repeats_equal_reference():
    """The high-repeat genome, 100 k pairs and 200 k single-end reads: ~15 % of the single-end reads tie between repeat
    copies, where the reference's rounded doubles decide between "unique" and "discarded" - the narrowed fp64 replay
    must reproduce every one."""
    out = _parity("cfg5", int(os.environ.get("PEMAP_CFG5_TEST_PAIRS", 100_000)), 40,
                  also_single=int(os.environ.get("PEMAP_CFG5_TEST_READS", 200_000)))
    assert out["single_end"]["cuda"]["replayed_fp64_read_mates"] > 0.05 * out["single_end"]["pairs"]
