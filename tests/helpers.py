"""Shared test helpers: seeded random taxonomies / hit tables for small parity cases.

Generators avoid, by construction, inputs on which the reference is non-deterministic (bean full
ties: same identifier under different ranks) unless a test asks for error cases explicitly."""
from __future__ import annotations

import json
import os
import random
import sys
from typing import Dict, List, Optional, Tuple

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

BACKBONE_RANKS = ["d", "p", "c", "o", "f", "g", "s"]
ODD_RANKS = ["clade", "no rank", "species group", "species subgroup", "strain", "subspecies", "k", "superkingdom",
             "Sub-Order", "u", "forma specialis", "serotype", "Domain", " Phylum "]


def random_taxonomy(rng: random.Random, n_leaves: int = 40, odd: float = 0.35, shared_root: bool = True,
                    truncate: float = 0.15) -> List[dict]:
    """A random tree; every leaf lineage becomes one taxonomy unit.  Identifiers are globally unique
    per node so that no two different (rank, identifier) pairs share an identifier."""
    counter = [0]

    def ident(prefix):
        counter[0] += 1
        return f"{prefix}{counter[0]}"

    # nodes at each backbone depth
    roots = [("d", ident("dom"))] if shared_root else [("d", ident("dom")) for _ in range(2)]
    units = []
    numid: Dict[str, int] = {}
    used_taxids = set()
    paths: List[List[Tuple[str, str]]] = [[r] for r in roots]
    for depth, rk in enumerate(BACKBONE_RANKS[1:], start=1):
        new_paths = []
        for p in paths:
            for _ in range(rng.choice([1, 1, 2, 3]) if len(paths) < n_leaves else 1):
                q = list(p)
                if rng.random() < odd / 2:
                    q.append((rng.choice(ODD_RANKS), ident("x")))
                q.append((rk, ident(rk)))
                new_paths.append(q)
        rng.shuffle(new_paths)
        paths = new_paths[: max(n_leaves, 1)]
    prefix_mode = rng.choice(["none"] * 6 + ["all"] * 3 + ["mixed"])
    prefix_rank = rng.choice(["no rank", "clade", "cellular root"])
    for p in paths:
        q = list(p)
        r = rng.random()
        if r < truncate:
            q = q[: rng.randint(1, len(q))]
        elif r < truncate + odd / 2:
            q.append((rng.choice(["strain", "subspecies", "no rank", "serotype"]), ident("t")))
        if prefix_mode == "all" or (prefix_mode == "mixed" and rng.random() < 0.5):
            q.insert(0, (prefix_rank, "cellular-organisms"))
        while True:
            t = rng.randint(1, 2_000_000_000)
            if t not in used_taxids:
                used_taxids.add(t)
                break
        text = ";".join(f"{a}__{b}" for a, b in q)
        num = ";".join(f"{a}__{numid.setdefault(b, 1000 + len(numid))}" for a, b in q)
        units.append({"taxid": t, "rank": q[-1][0], "numericLineage": num, "textLineage": text,
                      "accessions": [{"accession": f"NR_{t}.1", "oid": str(t)}]})
    return units


def taxonomy_json(units: List[dict]) -> dict:
    return {"blutilsVersion": "8.3.1", "ignoreTaxids": None, "replaceRank": None, "dropNonLinnaeanTaxonomies": False,
            "sourceDatabase": "/synthetic/db", "taxonomies": units}


def write_taxonomy(path: str, units: List[dict]) -> str:
    with open(path, "w") as f:
        json.dump(taxonomy_json(units), f)
    return path


def _fmt_pident(rng: random.Random, low: float = 60.0) -> str:
    v = rng.choice([100.0, 99.356, 98.927, 97.0, 96.999, 99.0, 92.0, 85.5, 80.001, 75.0, 60.0, 97.667, 98.333, 67.5])
    if rng.random() < 0.5:
        v = round(rng.uniform(low, 100.0), 3)
    s = f"{v:.3f}"
    if rng.random() < 0.3:
        s = s.rstrip("0").rstrip(".")
    return s


def random_blast(rng: random.Random, units: List[dict], n_queries: int = 30, max_hits: int = 12,
                 contiguous: bool = True, tie_rate: float = 0.6, low_pident: float = 60.0) -> bytes:
    taxids = [u["taxid"] for u in units]
    rows_by_q: List[List[str]] = []
    for qi in range(n_queries):
        q = rng.choice(["q%05d", "SRR1.%d_size_3", "draft-%d", "NR_%d.x"]) % qi
        nh = rng.randint(1, max_hits)
        base_bits = rng.choice([845, 833, 100, 84, 1200, 57])
        rows = []
        for h in range(nh):
            t = rng.choice(taxids[: max(3, len(taxids) // rng.choice([1, 2, 8]))]) if rng.random() < 0.7 else rng.choice(taxids)
            acc = f"NR_{rng.randint(100000, 100040)}.{rng.randint(1, 2)}"
            if rng.random() < tie_rate:
                bits = f"{base_bits}" if rng.random() < 0.7 else f"{base_bits}.{rng.randint(0, 9)}"
            else:
                bits = str(rng.randint(50, base_bits))
            if rng.random() < 0.05:
                bits = "%.3e" % float(bits)
            ln = rng.choice([455, 456, 1500, 200])
            rows.append("\t".join([q, acc, str(t), _fmt_pident(rng, low_pident), str(ln), str(rng.randint(0, 30)), str(rng.randint(0, 5)),
                                   "1", str(ln), str(rng.randint(1, 900)), str(rng.randint(901, 1800)),
                                   rng.choice(["0.0", "1e-50", "2.51e-117", "3.4", "1E-5", "5e+00"]), bits]))
        if rng.random() < 0.1 and rows:
            rows.append(rows[-1])  # exact duplicate row (exercises accessions dedup)
        rows_by_q.append(rows)
    if contiguous:
        lines = [r for rows in rows_by_q for r in rows]
    else:
        lines = [r for rows in rows_by_q for r in rows]
        # interleave: split some queries into several runs
        chunks = []
        for rows in rows_by_q:
            k = rng.randint(1, 3)
            cut = sorted(rng.sample(range(1, len(rows)), min(k - 1, max(0, len(rows) - 1)))) if len(rows) > 1 else []
            prev = 0
            for c in cut + [len(rows)]:
                chunks.append(rows[prev:c])
                prev = c
        rng.shuffle(chunks)
        # shuffling runs changes the relative order of a query's rows vs the file; that is fine (file order is what counts)
        lines = [r for ch in chunks for r in ch]
    text = "\n".join(lines)
    if rng.random() < 0.8:
        text += "\n"
    return text.encode()


def canon(results: List[dict]) -> List[dict]:
    """Sort by query bytes, drop runId."""
    out = []
    for r in results:
        r = dict(r)
        r.pop("runId", None)
        out.append(r)
    out.sort(key=lambda r: r["query"].encode())
    return out
