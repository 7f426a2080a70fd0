"""Synthesises hit tables from the reference's golden results (tests/golden/zymo_golden_derived.jsonl) so that the
golden file pins the CUDA path itself, not only the Python restatement.

The golden output holds results only (no BLAST table, no taxonomy DB).  Each result lists its consensus beans with
their lineage (`taxonomy`), `occurrences` and the number of (deduplicated) accessions, plus percIdentity / bitScore.
From that a top bit-score group is rebuilt: one lineage per bean (taxid = its index), one row per golden accession of the
bean (the golden accession strings themselves), all carrying the result's percIdentity and bitScore.  The accession
list of a bean is in the order of the sorted top group (lineage length, pident, align length, accession;
find_multi_taxa_consensus.rs:39-54); pident per row is not visible, so the rows of a bean get increasing align lengths
in the golden order -- the output must then list exactly the golden accessions in exactly the golden order.  The rows
are written to the table in a shuffled order, so the order really comes from the sort.
Which row the reference used as reference row is not visible in the output either, so every bean is tried as the
preferred reference (its rows get the align lengths that sort them first under `cautious` / last under `relaxed`,
find_multi_taxa_consensus.rs:39-63); a golden result is *pinned* when at least one variant reproduces every visible
field of it, accession lists included."""
import random
import json
import os

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zymo_golden_derived.jsonl")

FIELDS = ("reachedRank", "identifier", "percIdentity", "bitScore", "taxonomy", "singleMatch")
FULL_FIELDS = FIELDS + ("maxAllowedRank", "mutated")


def load():
    lines = open(FIX).read().splitlines()
    meta = json.loads(lines[0])
    return meta, [json.loads(l) for l in lines[1:]]


def build(strategy: str):
    """-> (taxids, lineages, text bytes, variants) with variants[query id] = (golden index, preferred bean)."""
    meta, golden = load()
    lin_id = {}
    for t in golden:
        for b in t["consensusBeans"]:
            lin_id.setdefault(b["taxonomy"], len(lin_id) + 1)
    rows = []
    variants = {}
    rng = random.Random(20261018)
    for gi, t in enumerate(golden):
        beans = t["consensusBeans"]
        pid = repr(float(t["percIdentity"]))
        bits = str(int(t["bitScore"])) if float(t["bitScore"]).is_integer() else repr(float(t["bitScore"]))
        for pref in range(len(beans)):
            q = f"g{gi:04d}_{pref:02d}"
            variants[q] = (gi, pref)
            qrows = []
            for bi, b in enumerate(beans):
                # the preferred bean's rows sort first (cautious: reference = first) / last (relaxed: reference = last) among
                # lineages of equal length; inside a bean the align length grows in the golden order of its accessions
                if strategy == "cautious":
                    base = 100 if bi == pref else 2000
                else:
                    base = 5000 if bi == pref else 2000
                assert b["occurrences"] == len(b["accessions"]) < 100
                for k, acc in enumerate(b["accessions"]):
                    aln = base + k
                    qrows.append(f"{q}\t{acc}\t{lin_id[b['taxonomy']]}\t{pid}\t{aln}\t0\t0\t1\t{aln}\t1\t{aln}\t0.0\t{bits}\n")
            rng.shuffle(qrows)
            rows += qrows
            # a lower-scoring hit that must not matter
            rows.append(f"{q}\tLOW.1\t{lin_id[beans[0]['taxonomy']]}\t80.0\t100\t0\t0\t1\t100\t1\t100\t0.0\t10\n")
    lineages = [None] * len(lin_id)
    for s, i in lin_id.items():
        lineages[i - 1] = s
    return list(range(1, len(lin_id) + 1)), lineages, "".join(rows).encode(), variants, golden, meta


def score(results, variants, golden):
    """results: list of {"query", "taxon"} dicts.  -> (pinned multi, pinned multi incl. maxAllowedRank/mutated, pinned single,
    total multi, total single), all weighted by the golden multiplicity."""
    ok = {}
    ok_full = {}
    for r in results:
        if r["query"] not in variants:
            continue
        gi, _ = variants[r["query"]]
        t, g = r["taxon"], golden[gi]
        same = t is not None and all(t[k] == g[k] for k in FIELDS)
        if same:
            gb = [(b["rank"], b["identifier"], b["occurrences"], b["taxonomy"], b["accessions"]) for b in g["consensusBeans"]]
            tb = [(b["rank"], b["identifier"], b["occurrences"], b["taxonomy"], b["accessions"]) for b in t["consensusBeans"]]
            same = gb == tb
        if same:
            ok[gi] = True
            if all(t[k] == g[k] for k in FULL_FIELDS):
                ok_full[gi] = True
    n_multi = sum(g["multiplicity"] for g in golden if not g["singleMatch"])
    n_single = sum(g["multiplicity"] for g in golden if g["singleMatch"])
    p_multi = sum(g["multiplicity"] for i, g in enumerate(golden) if not g["singleMatch"] and i in ok)
    p_full = sum(g["multiplicity"] for i, g in enumerate(golden) if not g["singleMatch"] and i in ok_full)
    p_single = sum(g["multiplicity"] for i, g in enumerate(golden) if g["singleMatch"] and i in ok_full)
    return p_multi, p_full, p_single, n_multi, n_single
