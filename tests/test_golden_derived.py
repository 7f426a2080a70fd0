"""Golden-derived known-answer test: the reference's own golden output
(test/mock/output/zymo-mock/blutils.consensus.json, reduced by tests/golden/make_golden_derived.py)
pins interpolate_identities / get_rank_adjusted_by_identity / get_adjusted_taxonomy_by_identity /
build_blast_consensus_identity / fold ordering.  The BLAST table + DB that produced it are not in the
reference repo, so each bean's lineage is tried as the reference lineage.

Expected (counted over the 2 283 results with a taxon): taxonomy+reachedRank+identifier reproduce for
2253/2253 multi-match results; the full tuple incl. maxAllowedRank/mutated for 1826 (the rest have a
Relaxed reference lineage that is not visible in the output); 30/30 single matches."""
import json
import os

import pyoracle as po
from pyoracle import _build

FIX = os.path.join(os.path.dirname(__file__), "golden", "zymo_golden_derived.jsonl")


def test_golden_derived():
    lines = open(FIX).read().splitlines()
    meta = json.loads(lines[0])
    assert meta["nResults"] == 3626 and meta["nNull"] == 1343 and meta["sortedByQuery"]
    bb = po.backbone_for(meta["taxon"])
    n = ok_tax = ok_full = single_n = single_ok = 0
    for ln in lines[1:]:
        t = json.loads(ln)
        mult = t["multiplicity"]
        beans = t["consensusBeans"]
        # bean order: occurrences desc, identifier asc (build_blast_consensus_identity.rs:50-60)
        keys = [(-b["occurrences"], b["identifier"].encode()) for b in beans]
        assert keys == sorted(keys)
        if t["singleMatch"]:
            single_n += mult
            b = beans[0]
            L = po.parse_lineage(b["taxonomy"])
            cut = po.interpolate([x[0] for x in L], bb)
            A = [L[k] for k in range(len(L)) if t["percIdentity"] >= cut[k]]
            good = (A and po.rank_full(A[-1][0]) == t["reachedRank"] and A[-1][1] == t["identifier"]
                    and po.lineage_str(A) == t["taxonomy"] and b["rank"] == t["reachedRank"] and b["identifier"] == t["identifier"]
                    and b["occurrences"] == 1 and t["maxAllowedRank"] is None and t["mutated"] is False)
            single_ok += mult if good else 0
            continue
        n += mult
        got_tax = got_full = False
        for b in beans:
            R = po.parse_lineage(b["taxonomy"])
            cut = po.interpolate([x[0] for x in R], bb)
            lv = [i for i, x in enumerate(R) if po.rank_full(x[0]) == b["rank"] and x[1] == b["identifier"]]
            if not lv:
                continue
            lv = lv[-1]
            single = len(beans) == 1
            idx = lv if single else lv - 1
            if idx < 0:
                continue
            fake = po.Row("x", 0, t["percIdentity"], 0, int(t["bitScore"]), 0)
            fb = []
            for b2 in beans:
                L2 = po.parse_lineage(b2["taxonomy"])
                for a in range(b2["nAccessions"]):
                    fb.append((L2[lv][0] if lv < len(L2) else R[lv][0], b2["identifier"], b2["taxonomy"], f"acc{a}"))
            out = _build(R, cut, bb, t["percIdentity"], single, idx, fb, fake)
            if (out["taxonomy"], out["reachedRank"], out["identifier"]) == (t["taxonomy"], t["reachedRank"], t["identifier"]):
                got_tax = True
                if out["maxAllowedRank"] == t["maxAllowedRank"] and out["mutated"] == t["mutated"]:
                    got_full = True
        ok_tax += mult if got_tax else 0
        ok_full += mult if got_full else 0
    assert (n, ok_tax, single_n, single_ok) == (2253, 2253, 30, 30)
    assert ok_full == 1826
