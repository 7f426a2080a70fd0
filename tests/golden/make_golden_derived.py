"""Generates tests/golden/zymo_golden_derived.jsonl from the reference's own golden output
(/root/reference/test/mock/output/zymo-mock/blutils.consensus.json, written by blutils 7.1.3).

The golden file holds results only (its BLAST table and taxonomy DB are not in the reference repo),
so it cannot be replayed end to end.  What it *does* pin is the arithmetic of
interpolate_identities / get_rank_adjusted_by_identity / get_adjusted_taxonomy_by_identity /
build_blast_consensus_identity / fold_consensus_list: each result is reduced to the fields those
functions produce (the accession list of every bean included: its order is the order of the sorted
top group, find_multi_taxa_consensus.rs:39-54), and identical reductions are de-duplicated with a
multiplicity.  Run here (the reference is not present on the GPU box):

    python tests/golden/make_golden_derived.py
"""
import json, collections, os

SRC = "/root/reference/test/mock/output/zymo-mock/blutils.consensus.json"
DST = os.path.join(os.path.dirname(__file__), "zymo_golden_derived.jsonl")

def main():
    d = json.load(open(SRC))
    seen = collections.OrderedDict()
    n_null = 0
    for r in d["results"]:
        t = r["taxon"]
        if t is None:
            n_null += 1
            continue
        red = {k: t[k] for k in ("reachedRank", "maxAllowedRank", "identifier", "percIdentity", "bitScore",
                                 "taxonomy", "mutated", "singleMatch")}
        red["consensusBeans"] = [
            {"rank": b["rank"], "identifier": b["identifier"], "occurrences": b["occurrences"],
             "taxonomy": b["taxonomy"], "nAccessions": len(b["accessions"]), "accessions": b["accessions"]}
            for b in t["consensusBeans"]]
        key = json.dumps(red, sort_keys=True)
        seen[key] = seen.get(key, 0) + 1
    with open(DST, "w") as f:
        f.write(json.dumps({"source": "test/mock/output/zymo-mock/blutils.consensus.json", "blutilsVersion":
                            d["config"]["blutilsVersion"], "taxon": d["config"]["taxon"], "nResults": len(d["results"]),
                            "nNull": n_null, "sortedByQuery": [r["query"] for r in d["results"]] ==
                            sorted(r["query"] for r in d["results"])}) + "\n")
        for k, mult in seen.items():
            o = json.loads(k); o["multiplicity"] = mult
            f.write(json.dumps(o) + "\n")
    print(len(seen), "unique of", len(d["results"]) - n_null, "->", DST, os.path.getsize(DST), "bytes")

if __name__ == "__main__":
    main()
