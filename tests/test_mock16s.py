"""BASELINE config C1: the reference's mock 16S database (tests/golden/mock16s, built by tests/golden/make_mock16s.py
from test/mock/input/ref_databases/mock-16S_taxonomies.tsv) + a hand-authored outfmt-6 table for the 10 mock query ids,
both strategies, text and numeric lineages, with the hit-less query coming from `headers`."""
import os

import pytest

import pyoracle as po
import sim_ffi
from oracle_ffi import Oracle, read_taxonomy_json

DIR = os.path.join(os.path.dirname(__file__), "golden", "mock16s")
TAX = os.path.join(DIR, "mock-16S.blutils.json")
TEXT = open(os.path.join(DIR, "blast.out"), "rb").read()
HEADERS = open(os.path.join(DIR, "headers.txt")).read().split()
CASES = [(u, s) for u in (True, False) for s in ("cautious", "relaxed")]


def expected(use_taxid, strategy):
    return open(os.path.join(DIR, f"expected.{'taxid' if use_taxid else 'text'}.{strategy}.jsonl"), "rb").read()


@pytest.mark.parametrize("use_taxid,strategy", CASES)
def test_mock16s_oracles_and_host_logic(use_taxid, strategy):
    want = expected(use_taxid, strategy)
    tax = po.load_taxonomy(TAX, use_taxid)
    assert po.results_to_jsonl(po.build_consensus_identities(TEXT, tax, "bacteria", strategy, headers=HEADERS)).encode() == want
    ids, lin = read_taxonomy_json(TAX, use_taxid)
    assert Oracle(ids, lin, "bacteria", strategy).run_raw(TEXT, headers=HEADERS)[0] == want
    assert sim_ffi.read_taxonomy_json(TAX, use_taxid) == len(ids) == 41
    rc, got, err = sim_ffi.run(ids, lin, "bacteria", strategy, TEXT, headers=HEADERS)
    assert rc == 0, err
    assert got == want


def test_mock16s_hand_checked_facts():
    """A few results checked by hand against the reference's rules."""
    import json

    res = {json.loads(l)["query"]: json.loads(l)["taxon"] for l in expected(True, "relaxed").decode().splitlines()}
    assert res["INVALID_SEQUENCE"] is None
    t = res["NR114924.257984.Bac"]  # single hit, 99.356 >= 99 -> species
    assert (t["singleMatch"], t["reachedRank"], t["identifier"], t["maxAllowedRank"], t["bitScore"]) == (True, "species", "257984", None, 845.0)
    t = res["draft-2582"]  # 81.25: d 60, clade 67.5, p 75, c 80 pass; o 85 fails
    assert (t["reachedRank"], t["taxonomy"]) == ("class", "d__2;clade__1783272;p__1239;c__91061")
    t = res["NR_113097.873513"]  # 84.2 / 84.9 truncate to one group of two strains of different species -> genus
    assert (t["bitScore"], t["reachedRank"], t["singleMatch"], len(t["consensusBeans"])) == (84.0, "genus", False, 2)
    t = res["NR025123.135626.Bac"]  # relaxed: all agree down to the shortest lineage, the longer reference goes deeper
    assert (t["reachedRank"], t["consensusBeans"][0]["rank"], t["consensusBeans"][0]["occurrences"]) == ("species", "family", 3)
    assert t["consensusBeans"][0]["accessions"] == ["NR025123.135626.Bacb", "NR025123.135626.Baca"]  # consecutive duplicate removed
    t = {json.loads(l)["query"]: json.loads(l)["taxon"] for l in expected(True, "cautious").decode().splitlines()}["NR025123.135626.Bac"]
    assert t["reachedRank"] == "family"  # cautious: the shortest lineage is the reference


@pytest.mark.gpu
@pytest.mark.parametrize("use_taxid,strategy", CASES)
def test_mock16s_gpu(use_taxid, strategy):
    from blutils_b200 import ConsensusStrategy, ParallelBlastOutput, Taxon, build_consensus_identities

    out = build_consensus_identities(ParallelBlastOutput(os.path.join(DIR, "blast.out"), HEADERS), TAX, Taxon.Bacteria,
                                     ConsensusStrategy.Cautious if strategy == "cautious" else ConsensusStrategy.Relaxed, use_taxid)
    assert out.jsonl() == expected(use_taxid, strategy)
    assert len(out) == 10 and out.n_rows == 31
