"""The reference's golden results (test/mock/output/zymo-mock/blutils.consensus.json, reduced to
tests/golden/zymo_golden_derived.jsonl) replayed END TO END: tests/golden_tables.py rebuilds a hit table + lineage map
from every golden result, the whole path (parse, join, group, top group, sort, level walk, bean fold, rank selection)
runs on it, and the output must reproduce the golden result's visible fields.

Counts (weighted by multiplicity; the same ones tests/test_golden_derived.py reaches through the arithmetic alone):
2253 / 2253 multi-match results incl. their beans (rank, identifier, occurrences, taxonomy, number of accessions), 1826 of
them incl. maxAllowedRank / mutated under `relaxed` (1596 under `cautious`; for the rest the reference row of the
original run is not visible in the output), 30 / 30 single matches.

CPU part: both oracles.  GPU part: the CUDA path through the C ABI -- which pins the kernels to reference-held data."""
import json

import pytest

import golden_tables as gt

EXPECT = {"cautious": (2253, 1596, 30, 2253, 30), "relaxed": (2253, 1826, 30, 2253, 30)}


@pytest.mark.parametrize("strategy", ["cautious", "relaxed"])
def test_oracles_reproduce_the_golden_results(strategy):
    import pyoracle as po
    from oracle_ffi import Oracle

    ids, lin, text, variants, golden, meta = gt.build(strategy)
    res = po.build_consensus_identities(text, dict(zip(ids, lin)), meta["taxon"], strategy)
    assert gt.score(res, variants, golden) == EXPECT[strategy]
    js = Oracle(ids, lin, meta["taxon"], strategy).run_raw(text)[0]
    assert js.decode() == po.results_to_jsonl(res)


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["cautious", "relaxed"])
def test_cuda_path_reproduces_the_golden_results(strategy):
    import pyoracle as po
    from blutils_b200 import ConsensusEngine, ConsensusStrategy, Taxon

    ids, lin, text, variants, golden, meta = gt.build(strategy)
    assert meta["taxon"] == "bacteria"
    eng = ConsensusEngine(Taxon.Bacteria, {"cautious": ConsensusStrategy.Cautious, "relaxed": ConsensusStrategy.Relaxed}[strategy])
    eng.load_taxonomy_arrays(ids, lin)
    out = eng.run_host(text)
    got = out.jsonl()
    res = [json.loads(l) for l in got.decode().splitlines()]
    assert gt.score(res, variants, golden) == EXPECT[strategy]
    # ... and every variant, pinned or not, is what the restatement says
    assert got.decode() == po.results_to_jsonl(po.build_consensus_identities(text, dict(zip(ids, lin)), "bacteria", strategy))
    out.close()
    eng.close()
