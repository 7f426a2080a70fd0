"""TEST INFRASTRUCTURE ONLY -- CPU restatement (Python mirror) of blutils 8.3.1's
consensus-identity path.  Nothing in the product (`blutils_b200/`) may import this.

PARITY STATUS: *parity unpinned* for the text-parsing / join / grouping layer (the
reference ships no tests and cannot be built here: no Rust toolchain, polars 0.37 /
serde_json / slugify 0.1 not vendored).  The interpolation + rank-selection + bean
folding arithmetic IS pinned against 2 283 real results of the reference's own golden
output `test/mock/output/zymo-mock/blutils.consensus.json` (tests/test_golden_derived.py)
and against the v8.3.1 example object in `docs/book/02_...md:192-248`.

Every function cites the reference file:line it restates (paths relative to
/root/reference).  This file is deliberately written independently of the C++ oracle
(oracle/blu_oracle.cpp); tests cross-check the two on random inputs.

Third-party behaviour restated from published semantics (not in /root/reference):
  * polars 0.37 CsvReader (core/Cargo.toml:27-31), call site mod.rs:357-371:
    '\n' rows, '\t' fields, no header, 13-column schema, ints/floats parsed
    correctly rounded; an unparsable or missing field aborts (`.finish().unwrap()` /
    `try_extract().unwrap()`, mod.rs:174-184,371).  Accepted grammar here (anything else is
    a DataError, i.e. "the reference aborts or its behaviour is unpinned"):
        int   := -?[0-9]{1,18}
        float := -?([0-9]+(\\.[0-9]*)?|\\.[0-9]+)([eE][+-]?[0-9]+)?   (correctly rounded)
    Empty lines are skipped; a '"' or '\r' byte anywhere is a DataError (polars quoting /
    CRLF handling not restated).
  * slugify 0.1 `slugify!` (linnaean_ranks.rs:69): ASCII only here.
  * serde_json 1.0 / ryu float formatting for the writer (write_blutils_output.rs).
"""
from __future__ import annotations

import json
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple


class DataError(Exception):
    """Inputs on which the reference panics or is unpinned."""


# --------------------------------------------------------------------------------------
# LinnaeanRank  (core/src/domain/dtos/linnaean_ranks.rs:14-107)
# A rank is a tuple: ("D", full_name) for the nine enum variants, ("O", slug) for Other.
# --------------------------------------------------------------------------------------
_FULL = ["undefined", "domain", "kingdom", "phylum", "class", "order", "family", "genus", "species"]
_BY_NAME = {n: ("D", n) for n in _FULL}
_BY_NAME.update({n[0]: ("D", n) for n in _FULL})  # "u","d","k","p","c","o","f","g","s"

Rank = Tuple[str, str]


def slugify(s: str) -> str:
    """slugify 0.1.0 `slugify!(s)` (sep '-', no stop words, no max length); ASCII input only."""
    if any(ord(c) > 127 for c in s):
        raise DataError("non-ASCII rank name (unidecode not restated)")
    s = s.lower().strip().strip("-").replace(" ", "-")
    out: List[str] = []
    is_sep = True
    for ch in s:
        if ("a" <= ch <= "z") or ("0" <= ch <= "9"):
            is_sep = False
            out.append(ch)
        elif not is_sep:
            is_sep = True
            out.append("-")
    if not out:
        raise DataError("empty rank slug (reference: slug.last().unwrap() panics)")
    if out[-1] == "-":
        out.pop()
    return "".join(out)


_RUST_WS = " \t\n\r\x0b\x0c"


def rank_from_str(s: str) -> Rank:
    """linnaean_ranks.rs:55-71."""
    t = s.lower().strip(_RUST_WS)
    if t in _BY_NAME:
        return _BY_NAME[t]
    return ("O", slugify(t))


def rank_display(r: Rank) -> str:
    """linnaean_ranks.rs:74-89 (`to_string`)."""
    return r[1][0] if r[0] == "D" else r[1]


def rank_full(r: Rank) -> str:
    """serde form (linnaean_ranks.rs:14-29) == as_full_rank_string (:92-106)."""
    return r[1]


# --------------------------------------------------------------------------------------
# Cutoff backbones (core/src/domain/dtos/taxon.rs:105-185) and CustomTaxon (:14-65)
# --------------------------------------------------------------------------------------
def _d(n: str) -> Rank:
    return ("D", n)


def backbone_for(taxon: str, custom: Optional[Dict[str, Optional[int]]] = None) -> List[Tuple[Rank, float]]:
    if taxon in ("fungi", "eukaryotes"):
        v = [97.0, 95.0, 90.0, 85.0, 80.0, 75.0, 60.0]
    elif taxon == "bacteria":
        v = [99.0, 97.0, 92.0, 85.0, 80.0, 75.0, 60.0]
    elif taxon == "custom":
        if custom is None:
            raise DataError("Custom taxon values are required")  # taxon.rs:117
        names = ["domain", "kingdom", "phylum", "class", "order", "family", "genus", "species"]
        if custom.get("domain") is None or custom.get("species") is None:
            raise DataError("custom cutoffs: domain and species are mandatory")  # taxon.rs:16-24
        return [(_d(n), float(custom.get(n) or 0)) for n in names]  # taxon.rs:123-139
    else:
        raise ValueError(taxon)
    names = ["species", "genus", "family", "order", "class", "phylum", "domain"]
    return [(_d(n), c) for n, c in zip(names, v)]


def load_custom_cutoffs(path: str) -> Dict[str, Optional[int]]:
    """CustomTaxon::from_file (taxon.rs:28-65): .yaml or .json chosen by extension, i16 fields."""
    import yaml

    if path.endswith(".yaml"):
        d = yaml.safe_load(open(path))
    elif path.endswith(".json"):
        d = json.load(open(path))
    else:
        raise DataError("Custom taxon file must be a YAML or JSON file")
    out = {}
    for k in ["domain", "kingdom", "phylum", "class", "order", "family", "genus", "species"]:
        v = d.get(k)
        if v is not None and not (isinstance(v, int) and -32768 <= v <= 32767):
            raise DataError("custom cutoff not an i16")
        out[k] = v
    return out


def rust_round3(v: float) -> float:
    """domain/utils/mod.rs:1-4 with decimals=3: (v*1000).round()/1000, half away from zero."""
    y = 1000.0
    x = v * y
    if math.isnan(x) or math.isinf(x):
        return x / y
    # f64::round = half away from zero; floor(|x|+0.5) is wrong just below .5, so compare the fraction
    fl = math.floor(abs(x))
    diff = abs(x) - fl
    r = fl + 1.0 if diff >= 0.5 else fl
    return math.copysign(r, x) / y


def interpolate(ranks: Sequence[Rank], backbone: List[Tuple[Rank, float]]) -> List[float]:
    """InterpolatedIdentity::interpolate_identities (linnaean_ranks.rs:220-383).

    Elements are ('def', rank, cutoff) / ('non', display_string, cutoff); equality is
    tuple equality like the derived PartialEq on RankedLinnaeanIdentity."""
    m = []
    for r in ranks:  # :239-261
        hit = next((b for b in backbone if b[0] == r), None)
        m.append(("def", hit[0], hit[1]) if hit is not None else ("non", rank_display(r), 0.0))
    if all(e[0] == "def" for e in m):  # :265-270
        return [e[2] for e in m]
    updated: Dict[int, float] = {}
    for n, e in enumerate(m):  # :275-333
        if e[0] != "non":
            continue
        previous = next((x for x in reversed(m[:n]) if x[0] == "def"), m[0])
        p = next((i for i, x in enumerate(m) if x == previous), 0)
        nxt = next((x for x in m[n:] if x[0] == "def"), m[-1])
        q = next((i for i, x in enumerate(m) if x == nxt), len(m) - 1)
        window = m[p:][: q + 1]  # skip_while(!= previous).take(next_index+1)  :324-329
        t = n - p
        first = window[0][2] if window[0][0] == "def" else backbone[0][1]  # :341-347
        last = window[-1][2] if window[-1][0] == "def" else 100.0  # :349-353
        weight = last - first
        size = float(len(window) - 1)
        try:
            step = weight / size
        except ZeroDivisionError:  # IEEE: x/0.0
            step = math.nan if weight == 0.0 else math.copysign(math.inf, weight)
        updated[n] = rust_round3(first + (float(t) * step))  # :358-362
    return [updated.get(i, 100.0) if e[0] == "non" else e[2] for i, e in enumerate(m)]


# --------------------------------------------------------------------------------------
# Lineage parsing (core/src/domain/dtos/blast_result.rs:38-120)
# --------------------------------------------------------------------------------------
def parse_lineage(s: str) -> List[Tuple[Rank, str]]:
    out = []
    for part in s.split(";"):
        pieces = part.split("__")
        if len(pieces) != 2:
            raise DataError("Unexpected error on parse taxonomy")  # :109-114 -> panic at fsqc.rs:59
        out.append((rank_from_str(pieces[0]), pieces[1]))
    return out


def bean_str(b: Tuple[Rank, str]) -> str:
    """TaxonomyBean::taxonomy_to_string (taxonomy_bean.rs:19-27)."""
    return f"{rank_display(b[0])}__{b[1]}"


def lineage_str(L: Sequence[Tuple[Rank, str]]) -> str:
    return ";".join(bean_str(b) for b in L)


# --------------------------------------------------------------------------------------
# Taxonomy file (mod.rs:246-327; taxonomies_map.rs:6-32)
# --------------------------------------------------------------------------------------
def load_taxonomy(path: str, use_taxid: bool) -> Dict[int, str]:
    d = json.load(open(path))
    for k in ("blutilsVersion", "sourceDatabase", "taxonomies"):
        if k not in d:
            raise IOError(f"taxonomies json: missing {k}")
    out: Dict[int, str] = {}
    for u in d["taxonomies"]:
        for k in ("taxid", "rank", "numericLineage", "textLineage", "accessions"):
            if k not in u:
                raise IOError(f"taxonomies json: missing {k}")
        tid = u["taxid"]
        if not isinstance(tid, int) or tid < 0 or tid >= 2**64:
            raise IOError("taxid not u64")
        f = float(tid)  # to_f64() mod.rs:278, then cast Int64 :309
        if f >= 9223372036854775808.0:
            continue  # cast overflow -> null key, never joins
        key = int(f)
        if key in out:
            raise DataError("duplicate taxid in taxonomy (left join would duplicate rows; unsupported)")
        out[key] = u["numericLineage"] if use_taxid else u["textLineage"]
    return out


# --------------------------------------------------------------------------------------
# outfmt-6 parsing (mod.rs:226-244,329-373 + fold_results_by_query :134-221)
# --------------------------------------------------------------------------------------
_INT = re.compile(rb"-?[0-9]{1,18}\Z")
_FLT = re.compile(rb"(-?)(?:([0-9]+)(?:\.([0-9]*))?|\.([0-9]+))(?:[eE]([+-]?[0-9]+))?\Z")


def parse_int(b: bytes) -> int:
    if not _INT.match(b):
        raise DataError(f"bad integer field {b!r}")
    return int(b)


def check_float(b: bytes) -> None:
    if not _FLT.match(b):
        raise DataError(f"bad float field {b!r}")


def parse_float(b: bytes) -> float:
    """Correctly rounded decimal->f64 (the CUDA path reports BLU_ERR_UNSUPPORTED outside its exact range)."""
    mt = _FLT.match(b)
    if not mt:
        raise DataError(f"bad float field {b!r}")
    return float(b)  # Python float(): correctly rounded, any number of digits


@dataclass
class Row:
    acc: str
    taxid: int
    pident: float
    alnlen: int
    bits: int
    order: int  # file order


def parse_blast(text: bytes) -> Dict[str, List[Row]]:
    if len(text) == 0:
        raise DataError("empty blast output (polars: empty CSV)")
    if b'"' in text or b"\r" in text:
        raise DataError("quote or CR byte in blast output (unsupported)")
    groups: Dict[str, List[Row]] = {}
    n = 0
    for line in text.split(b"\n"):
        if not line:
            continue
        f = line.split(b"\t")
        if len(f) != 13:
            raise DataError("row does not have 13 fields")
        if len(f[0]) == 0 or len(f[1]) == 0:
            raise DataError("empty query/accession field (null string; unsupported)")
        taxid = parse_int(f[2])
        pident = parse_float(f[3])
        alnlen = parse_int(f[4])
        for k in range(5, 11):
            parse_int(f[k])
        check_float(f[11])
        bs = parse_float(f[12])
        if not (abs(bs) < 9223372036854775808.0):
            raise DataError("bit score out of i64 range")
        bits = int(bs)  # truncation toward zero: try_extract::<i64> mod.rs:184
        try:
            q = f[0].decode("utf-8")
            acc = f[1].decode("utf-8")
        except UnicodeDecodeError:
            raise DataError("invalid utf-8")
        groups.setdefault(q, []).append(Row(acc, taxid, pident, alnlen, bits, n))
        n += 1
    if n == 0:
        raise DataError("no rows")
    return groups


# --------------------------------------------------------------------------------------
# Consensus (find_single_query_consensus.rs, find_multi_taxa_consensus.rs,
# build_blast_consensus_identity.rs, consensus_result.rs:48-88)
# --------------------------------------------------------------------------------------
def bean_key_order(tax: Dict[int, str]) -> Dict[str, int]:
    """Order of first appearance of every "{rank}__{identifier}" key in the taxonomy map (lineages in map order, levels root to
    leaf): the tie-break of _fold_beans."""
    order: Dict[str, int] = {}
    for s in tax.values():
        try:
            L = parse_lineage(s)
        except DataError:
            continue
        for rank, ident in L:
            order.setdefault(f"{rank_display(rank)}__{ident}", len(order))
    return order


def _fold_beans(beans: List[Tuple[Rank, str, str, str]], key_order: Optional[Dict[str, int]] = None) -> List[dict]:
    """ConsensusBean::fold_consensus_list (consensus_result.rs:65-88) + the sort at
    build_blast_consensus_identity.rs:50-60.  beans: (rank, identifier, taxonomy, accession)."""
    acc: Dict[str, dict] = {}
    for rank, ident, tax, a in beans:
        key = f"{rank_display(rank)}__{ident}"
        b = acc.get(key)
        if b is None:
            b = acc[key] = {"rank": rank_full(rank), "identifier": ident, "occurrences": 0, "taxonomy": tax, "accessions": []}
        b["accessions"].append(a)
        d = []
        for x in b["accessions"]:  # Vec::dedup -- consecutive only
            if not d or d[-1] != x:
                d.append(x)
        b["accessions"] = d
        b["occurrences"] += 1
    # Full ties (same occurrences and identifier, different rank) leave the reference in HashMap order: non-deterministic
    # there.  Every implementation of this repository breaks them the same way: by the order in which the beans' keys first
    # appear in the taxonomy map (the product's dictionary ids, blu_taxonomy.cpp); without a map, first-seen order.
    keys = list(acc.keys())
    out = sorted(range(len(keys)), key=lambda i: (-acc[keys[i]]["occurrences"], acc[keys[i]]["identifier"].encode("utf-8"),
                                                  key_order.get(keys[i], 1 << 60) if key_order is not None else 0, i))
    return [acc[keys[i]] for i in out]


def _allowed(cut: List[float], identity: float) -> Optional[int]:
    """get_rank_adjusted_by_identity (linnaean_ranks.rs:174-192): first j with !(identity > cut[j])."""
    for j, c in enumerate(cut):
        if not (identity > c):
            return j
    return None


def _allowed_rank(ranks: Sequence[Rank], backbone, j: int) -> Rank:
    """build_blast_consensus_identity.rs:22-30: DefaultRank -> rank, NonDefaultRank(s) -> Other(s)."""
    r = ranks[j]
    if any(b[0] == r for b in backbone):
        return r
    return ("O", rank_display(r))


def _build(R, cut, backbone, identity, single, idx, beans, ref_row: Row, key_order: Optional[Dict[str, int]] = None) -> dict:
    """build_blast_consensus_identity (build_blast_consensus_identity.rs:9-105)."""
    ranks = [b[0] for b in R]
    bean = R[idx]
    j = _allowed(cut, identity)
    max_allowed = None
    mutated = False
    if j is not None:
        ar = _allowed_rank(ranks, backbone, j)
        max_allowed = rank_full(ar)
        mutated = bean[0] != ar  # :35-37
    folded = _fold_beans(beans, key_order)
    F = [R[k] for k in range(len(R)) if identity >= cut[k]]  # linnaean_ranks.rs:194-212
    if single and len(folded) == 1:
        A = F
    else:
        A = F[: idx + 1]  # enumerate().take_while(index <= bean_index) AFTER the filter :72-82
    last = A[-1] if A else R[idx]
    return {
        "reachedRank": rank_full(last[0]),
        "maxAllowedRank": max_allowed,
        "identifier": last[1],
        "percIdentity": ref_row.pident,
        "bitScore": float(ref_row.bits),
        "taxonomy": lineage_str(A),
        "mutated": mutated,
        "singleMatch": False,
        "consensusBeans": folded,
    }


def consensus_for_query(rows: List[Row], tax: Dict[int, str], backbone, strategy: str,
                        lineage_cache: Optional[dict] = None, key_order: Optional[Dict[str, int]] = None) -> dict:
    """find_single_query_consensus.rs:17-173 -> taxon object (dict in serde field order)."""
    top = max(r.bits for r in rows)  # :28-50, only the first loop iteration is reachable
    G = [r for r in rows if r.bits == top]
    Ls = []
    for r in G:
        s = tax.get(r.taxid)
        if s is None:
            raise DataError("unmapped taxid in top bit-score group (lineage 'null')")  # mod.rs:185 + fsqc.rs:59
        if lineage_cache is not None:
            L = lineage_cache.get(s)
            if L is None:
                L = lineage_cache[s] = parse_lineage(s)
        else:
            L = parse_lineage(s)
        Ls.append(L)
    if len(G) == 1:  # :74-150
        r, L = G[0], Ls[0]
        cut = interpolate([b[0] for b in L], backbone)
        A = [L[k] for k in range(len(L)) if r.pident >= cut[k]]
        if not A:
            raise DataError("No taxonomy found for result")  # :113-119
        last = A[-1]
        return {
            "reachedRank": rank_full(last[0]),
            "maxAllowedRank": None,
            "identifier": last[1],
            "percIdentity": r.pident,
            "bitScore": float(r.bits),
            "taxonomy": lineage_str(A),
            "mutated": False,
            "singleMatch": True,
            "consensusBeans": _fold_beans([(last[0], last[1], lineage_str(L), r.acc)]),
        }
    # find_multi_taxa_consensus.rs:22-217
    order = sorted(range(len(G)), key=lambda i: (len(Ls[i]), G[i].pident, G[i].alnlen, G[i].acc.encode("utf-8"), i))
    S = [G[i] for i in order]
    SL = [Ls[i] for i in order]
    ref_i = 0 if strategy == "cautious" else len(S) - 1  # :60-63
    R = SL[ref_i]
    ref_row = S[ref_i]
    cut = interpolate([b[0] for b in R], backbone)
    out = None
    for i in range(len(R)):  # :137-214
        level = [(L, r) for L, r in zip(SL, S)]
        # take_while(index < taxonomy.len()) over a length-ascending list
        tw = []
        for L, r in level:
            if i < len(L):
                tw.append((L, r))
            else:
                break
        keys = {rank_display(L[i][0]) + L[i][1] for L, _ in tw}  # no separator :153-157
        if not keys:
            continue
        beans = [(L[i][0], L[i][1], lineage_str(L), r.acc) for L, r in tw]
        if len(keys) > 1:
            if i == 0:
                raise DataError("root-level disagreement (index - 1 underflow)")  # :181
            mx = 0.0
            for _, r in tw:  # fold from 0.0 :182-185
                if r.pident > mx:
                    mx = r.pident
            out = _build(R, cut, backbone, mx, False, i - 1, beans, ref_row, key_order)
            break
        out = _build(R, cut, backbone, ref_row.pident, True, i, beans, ref_row, key_order)
    assert out is not None
    return out


def build_consensus_identities(text: bytes, tax: Dict[int, str], taxon: str, strategy: str,
                               custom: Optional[dict] = None, headers: Optional[List[str]] = None) -> List[dict]:
    """build_consensus_identities (mod.rs:40-129) followed by the writer's flatten + sort
    (write_blutils_output.rs:87-111).  Returns [{"query":..., "taxon": {...}|None}] sorted by query."""
    backbone = backbone_for(taxon, custom)
    groups = parse_blast(text)
    cache: dict = {}
    key_order = bean_key_order(tax)
    res = []
    for q, rows in groups.items():
        res.append({"query": q, "taxon": consensus_for_query(rows, tax, backbone, strategy, cache, key_order)})
    if headers:
        for h in headers:  # mod.rs:91-100
            if h not in groups:
                res.append({"query": h, "taxon": None})
    res.sort(key=lambda r: r["query"].encode("utf-8"))
    return res


# --------------------------------------------------------------------------------------
# serde_json-compatible serialisation (write_blutils_output.rs:126-215)
# --------------------------------------------------------------------------------------
def ryu_f64(v: float) -> str:
    """serde_json float formatting = ryu `format64` pretty printer (finite values)."""
    if v != v or v in (math.inf, -math.inf):
        return "null"  # serde_json writes null for non-finite
    if v == 0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    r = repr(abs(v))  # shortest round-trip digits
    if "e" in r:
        mant, ex = r.split("e")
        ex = int(ex)
    else:
        mant, ex = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp)
    k = ex - len(fp)
    stripped = digits.lstrip("0")
    digits = stripped
    t = digits.rstrip("0")
    k += len(digits) - len(t)
    digits = t
    length = len(digits)
    kk = length + k
    sign = "-" if v < 0 else ""
    if 0 <= k and kk <= 16:
        s = digits + "0" * k + ".0"
    elif 0 < kk <= 16:
        s = digits[:kk] + "." + digits[kk:]
    elif -5 < kk <= 0:
        s = "0." + "0" * (-kk) + digits
    elif length == 1:
        s = f"{digits}e{kk - 1}"
    else:
        s = f"{digits[0]}.{digits[1:]}e{kk - 1}"
    return sign + s


def _jstr(s: str) -> str:
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif o < 0x20:
            out.append({8: "\\b", 9: "\\t", 10: "\\n", 12: "\\f", 13: "\\r"}.get(o, "\\u%04x" % o))
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def to_json_compact(x) -> str:
    if x is None:
        return "null"
    if x is True:
        return "true"
    if x is False:
        return "false"
    if isinstance(x, float):
        return ryu_f64(x)
    if isinstance(x, int):
        return str(x)
    if isinstance(x, str):
        return _jstr(x)
    if isinstance(x, list):
        return "[" + ",".join(to_json_compact(v) for v in x) + "]"
    if isinstance(x, dict):
        return "{" + ",".join(_jstr(k) + ":" + to_json_compact(v) for k, v in x.items()) + "}"
    raise TypeError(type(x))


def to_json_pretty(x, ind: int = 0) -> str:
    pad = "  " * (ind + 1)
    if isinstance(x, list):
        if not x:
            return "[]"
        return "[\n" + ",\n".join(pad + to_json_pretty(v, ind + 1) for v in x) + "\n" + "  " * ind + "]"
    if isinstance(x, dict):
        if not x:
            return "{}"
        return "{\n" + ",\n".join(pad + _jstr(k) + ": " + to_json_pretty(v, ind + 1) for k, v in x.items()) + "\n" + "  " * ind + "}"
    return to_json_compact(x)


def results_to_jsonl(results: List[dict], run_id: Optional[str] = None) -> str:
    """JSONL body lines (without the leading config line); runId omitted when None (masked)."""
    lines = []
    for r in results:
        o = {}
        if run_id is not None:
            o["runId"] = run_id
        o["query"] = r["query"]
        o["taxon"] = r["taxon"]
        lines.append(to_json_compact(o))
    return "\n".join(lines) + ("\n" if lines else "")


# --------------------------------------------------------------------------------------
# parse_consensus_as_tabular (core/src/use_cases/parse_consensus_as_tabular/mod.rs:15-173)
# --------------------------------------------------------------------------------------
def rust_display_f64(v: float) -> str:
    """Rust `format!("{}", f64)`: shortest round-trip digits, positional, no trailing '.0'."""
    from decimal import Decimal

    if v != v:
        return "NaN"
    if v in (math.inf, -math.inf):
        return "inf" if v > 0 else "-inf"
    s = format(Decimal(repr(v)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s


def results_to_tabular(results: List[dict], run_id: str, to_stdout: bool) -> str:
    """The byte stream the reference produces: println! per piece on stdout, raw pieces in a file (mod.rs:58-170,
    shared/write_file_or_stdout.rs:3-22)."""
    end = "\n" if to_stdout else ""
    out = ["\t".join(["run-id", "query", "type", "rank", "identifier", "perc-identity", "bit-score", "taxonomy", "mutated",
                      "single-match", "occurrences", "accessions"]) + end]
    for r in results:
        t = r["taxon"]
        if t is None:
            out.append(f"{r['query']}\tnull\n" + end)
            continue
        b = "true" if t["mutated"] else "false"
        sm = "true" if t["singleMatch"] else "false"
        out.append("\t".join([run_id, r["query"], "consensus", t["reachedRank"], t["identifier"], rust_display_f64(t["percIdentity"]),
                              rust_display_f64(t["bitScore"]), t["taxonomy"] if t["taxonomy"] is not None else "null", b, sm, "null",
                              "null"]) + end)
        for c in t["consensusBeans"] or []:
            out.append("\t".join([run_id, r["query"], "blast-match", c["rank"], c["identifier"], "null", rust_display_f64(t["bitScore"]),
                                  c["taxonomy"] if c["taxonomy"] is not None else "null", "null", "null", str(c["occurrences"]),
                                  ", ".join(c["accessions"])]) + end)
    return "".join(out)


# --------------------------------------------------------------------------------------
# serde_yaml 0.9 serialisation (write_blutils_output.rs:216-248: serde_yaml::to_writer of BlutilsOutput)
#
# serde_yaml is not vendored under /root/reference (core/Cargo.toml: serde_yaml = "0.9"): PARITY UNPINNED, restated from
# the crate's published behaviour:
#   * block style; a sequence that is a mapping value starts at the key's indentation ("- " under the key);
#   * f64 through ryu (same text as serde_json), bool / null / integers plain;
#   * a string is single-quoted when it would read back as another type (ser.rs serialize_str -> de.rs
#     visit_untagged_scalar: empty, null / ~, true / false, integers incl. 0x / 0o / 0b, floats incl. .inf / .nan) or is one
#     of the YAML 1.1 booleans (y, yes, n, no, on, off: "ambiguous" strings);
#   * otherwise libyaml's emitter decides (yaml_emitter_analyze_scalar / select_scalar_style, block context): plain unless
#     the string has a leading / trailing space, a non-printable character, starts with an indicator
#     (# , [ ] { } & * ! | > ' " % @ `), with "- " / "? " / ": " (or is just that character), with "---" / "...", or contains
#     ": " / a trailing ":" / " #"; then single-quoted ('' escapes '), double-quoted only for non-printable characters.
# --------------------------------------------------------------------------------------
_YAML11_BOOLS = {"y", "yes", "n", "no", "on", "off", "true", "false", "null", "~"}
_FLOAT_RE = re.compile(r"[+-]?([0-9]+(\.[0-9]*)?|\.[0-9]+)([eE][+-]?[0-9]+)?\Z")


def _digits_but_not_number(s: str) -> bool:
    t = s[1:] if s[:1] in "+-" and s else s
    return len(t) > 1 and t[0] == "0" and all("0" <= c <= "9" for c in t[1:])


def _yaml_int_like(s: str) -> bool:
    body = s[1:] if s[:1] in ("+", "-") else s
    if not body or body[:1] in ("+", "-"):
        return False
    for pre, digits in (("0x", "0123456789abcdefABCDEF"), ("0o", "01234567"), ("0b", "01")):
        if body.startswith(pre):
            rest = body[2:]
            return bool(rest) and all(c in digits for c in rest) and int(rest, {"0x": 16, "0o": 8, "0b": 2}[pre]) < (1 << 128)
    if _digits_but_not_number(s):
        return False
    return all("0" <= c <= "9" for c in body) and int(body) < (1 << 128)


def _yaml_float_like(s: str) -> bool:
    if _digits_but_not_number(s):
        return False
    u = s[1:] if s.startswith("+") else s
    if s.startswith("+") and u[:1] in ("+", "-"):
        return False
    if u in (".inf", ".Inf", ".INF") or s in ("-.inf", "-.Inf", "-.INF", ".nan", ".NaN", ".NAN"):
        return True
    if not _FLOAT_RE.match(u):
        return False
    try:
        return math.isfinite(float(u))
    except (ValueError, OverflowError):
        return False


def _yaml_printable(cp: int) -> bool:
    return (cp == 0x0A or 0x20 <= cp <= 0x7E or cp == 0x85 or 0xA0 <= cp <= 0xD7FF or (0xE000 <= cp <= 0xFFFD and cp != 0xFEFF)
            or 0x10000 <= cp <= 0x10FFFF)


def yaml_str(s: str) -> str:
    if s == "" or s in ("null", "Null", "NULL", "~", "true", "True", "TRUE", "false", "False", "FALSE") or _yaml_int_like(s) or _yaml_float_like(s) \
            or s.lower() in _YAML11_BOOLS:
        return "'" + s.replace("'", "''") + "'"
    special = any(not _yaml_printable(ord(c)) or c == "\n" for c in s)
    if special:
        esc = {0: "\\0", 7: "\\a", 8: "\\b", 9: "\\t", 10: "\\n", 11: "\\v", 12: "\\f", 13: "\\r", 27: "\\e", 34: '\\"', 92: "\\\\", 0x85: "\\N",
               0xA0: "\\_", 0x2028: "\\L", 0x2029: "\\P"}
        out = ['"']
        for c in s:
            cp = ord(c)
            if not _yaml_printable(cp) or cp in (0xFEFF, 0x0A, 0x0D, 0x85, 0x2028, 0x2029, 34, 92):
                if cp in esc:
                    out.append(esc[cp])
                elif cp <= 0xFF:
                    out.append("\\x%02X" % cp)
                elif cp <= 0xFFFF:
                    out.append("\\u%04X" % cp)
                else:
                    out.append("\\U%08X" % cp)
            else:
                out.append(c)
        out.append('"')
        return "".join(out)
    block_ind = s.startswith("---") or s.startswith("...")
    n = len(s)
    for i, c in enumerate(s):
        nxt_blank = i + 1 >= n or s[i + 1] in " \t"
        if i == 0:
            if c in "#,[]{}&*!|>'\"%@`":
                block_ind = True
            if c in "?:" and nxt_blank:
                block_ind = True
            if c == "-" and nxt_blank:
                block_ind = True
        else:
            if c == ":" and nxt_blank:
                block_ind = True
            if c == "#" and s[i - 1] in " \t":
                block_ind = True
    if block_ind or s[0] == " " or s[-1] == " ":
        return "'" + s.replace("'", "''") + "'"
    return s


def results_to_yaml(results: List[dict], run_id: str) -> str:
    """serde_yaml::to_writer(BlutilsOutput { results, config: None }) (write_blutils_output.rs:216-248)."""
    o: List[str] = []
    o.append("results: []\n" if not results else "results:\n")
    for r in results:
        o.append("- runId: " + yaml_str(run_id) + "\n  query: " + yaml_str(r["query"]) + "\n")
        t = r["taxon"]
        if t is None:
            o.append("  taxon: null\n")
            continue
        o.append("  taxon:\n    reachedRank: " + yaml_str(t["reachedRank"]) + "\n    maxAllowedRank: " +
                 ("null" if t["maxAllowedRank"] is None else yaml_str(t["maxAllowedRank"])) + "\n    identifier: " + yaml_str(t["identifier"]) +
                 "\n    percIdentity: " + ryu_f64(t["percIdentity"]) + "\n    bitScore: " + ryu_f64(t["bitScore"]) + "\n    taxonomy: " +
                 yaml_str(t["taxonomy"]) + "\n    mutated: " + ("true" if t["mutated"] else "false") + "\n    singleMatch: " +
                 ("true" if t["singleMatch"] else "false") + "\n")
        beans = t["consensusBeans"]
        o.append("    consensusBeans: []\n" if not beans else "    consensusBeans:\n")
        for b in beans:
            o.append("    - rank: " + yaml_str(b["rank"]) + "\n      identifier: " + yaml_str(b["identifier"]) + "\n      occurrences: " +
                     str(b["occurrences"]) + "\n      taxonomy: " + yaml_str(b["taxonomy"]) + "\n")
            o.append("      accessions: []\n" if not b["accessions"] else "      accessions:\n")
            for a in b["accessions"]:
                o.append("      - " + yaml_str(a) + "\n")
    o.append("config: null\n")
    return "".join(o)
