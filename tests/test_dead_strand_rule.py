"""CPU: the seed kernel's dead-strand rule (seed_rbi.cuh: with the running min_match at F the first nseg - F + 2
segments decide whether the strand can change the hit list) against the full find_matches restatement
(pemapper.c:2189-2288, tests/seed_model.py) on adversarial genomes, with and without synthetic chance hits, at two
values of max_hits.  Whenever the rule declares a strand dead, skipping it must not change initial_map's result."""
import numpy as np

import oracle_lib as ol
import seed_model as model


def test_dead_strands_change_nothing():
    for kw in (dict(n_reads=220, seed=41), dict(n_reads=160, seed=42, noise=0.7, n_random=120_000)):
        rep = model.main(max_hits_list=(200, 12), **kw)
        for max_hits, r in rep.items():
            assert r["mismatches"] == 0, (kw, max_hits, r)
        assert rep[200]["dead"] > rep[200]["reads"] // 5, (kw, rep)   # the rule does fire (after a full-length hit)


def test_model_equals_pinned_oracle(oracle_built):
    """The model's initial_map (no rule) is the pinned C oracle's, candidate list for candidate list, in order."""
    rng = np.random.default_rng(43)
    genome = model.make_genome(rng, 150_000)
    ix = model.Index(genome)
    reads = model.make_reads(rng, genome, 150)
    orc = ol.Oracle([genome])
    st = dict(strands=0, dead=0)
    for r in reads:
        spots, orients = orc.initial_map(r)
        want = list(zip((int(x) for x in spots), (int(x) for x in orients)))
        assert model.initial_map(ix, r, 200, False, st) == want, r
    orc.close()
