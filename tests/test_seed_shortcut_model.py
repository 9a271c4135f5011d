"""CPU: the single-chain shortcut of the seed kernel (seed_rbi.cuh, try_single_chain), as modelled in
tools/seed_shortcut_model.py, against the full find_matches restatement (pemapper.c:2189-2288) on adversarial genomes:
tandem repeats, diverged copies, crowded k-mers, low complexity, reads of both strands with substitutions, indels
and N.  Whenever the shortcut makes a claim it must be the full computation's result, at three values of max_hits."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import seed_shortcut_model as model  # noqa: E402


def test_shortcut_claims_equal_full_rule():
    rep = model.main(n_reads=250, seed=21, max_hits_list=(200, 12))
    for max_hits, r in rep.items():
        assert r["mismatches"] == 0, (max_hits, r)
        assert r["claims"] > r["reads"] // 2, (max_hits, r)   # the shortcut does fire on ordinary reads
