"""`blu blastn build-tabular`: blutils result file -> 12-column TSV, through the C ABI
(`blu_result_file_to_tabular`, no GPU needed) and the CLI shim.  Expected bytes come from the Python oracle's restatement of
parse_consensus_as_tabular (mod.rs:15-173); the input files are what write_blutils_output produces for the mock 16S run
(serde_json pretty / compact / JSONL)."""
import json
import os
import subprocess

import pytest

import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DIR = os.path.join(HERE, "golden", "mock16s")
RUN_ID = "0b0e3c55-7a2f-4a61-9d4e-5f1c2a7b8c9d"
OTHER_ID = "11111111-2222-4333-8444-555555555555"


def _results(use_taxid=False, strategy="relaxed"):
    text = open(os.path.join(DIR, "blast.out"), "rb").read()
    headers = open(os.path.join(DIR, "headers.txt")).read().split()
    tax = po.load_taxonomy(os.path.join(DIR, "mock-16S.blutils.json"), use_taxid)
    return po.build_consensus_identities(text, tax, "bacteria", strategy, headers=headers)


def _doc(results, run_id=RUN_ID, config=None):
    return {"results": [dict(([("runId", run_id)] if run_id else []) + list(r.items())) for r in results], "config": config}


def _tab(in_path, out_path, fmt="json", run_id=None):
    from blutils_b200 import OutputFormat, parse_consensus_as_tabular

    parse_consensus_as_tabular(in_path, out_path, {"json": OutputFormat.Json, "jsonl": OutputFormat.Jsonl, "yaml": OutputFormat.Yaml}[fmt], run_id)


@pytest.mark.parametrize("use_taxid,strategy", [(False, "relaxed"), (True, "cautious")])
@pytest.mark.parametrize("pretty", [True, False])
def test_json_to_tsv_file(tmp_path, use_taxid, strategy, pretty):
    res = _results(use_taxid, strategy)
    src = tmp_path / "blutils.json"
    src.write_text(po.to_json_pretty(_doc(res)) if pretty else po.to_json_compact(_doc(res)))
    _tab(str(src), str(tmp_path / "table.txt"))  # extension forced to .tsv
    want = po.results_to_tabular(res, RUN_ID, to_stdout=False)
    assert (tmp_path / "table.tsv").read_text() == want
    assert any(r["taxon"] is None for r in res) and any(r["taxon"] and r["taxon"]["consensusBeans"] for r in res)
    # an existing output is replaced, not appended to
    _tab(str(src), str(tmp_path / "table.tsv"))
    assert (tmp_path / "table.tsv").read_text() == want


def test_run_id_sources(tmp_path):
    res = _results()
    # no per-result runId: the config's run id is used ...
    src = tmp_path / "a.json"
    src.write_text(po.to_json_compact(_doc(res, run_id=None, config={"runId": OTHER_ID, "isConfig": True, "anything": [1, {"x": None}]})))
    _tab(str(src), str(tmp_path / "a.tsv"))
    assert (tmp_path / "a.tsv").read_text() == po.results_to_tabular(res, OTHER_ID, to_stdout=False)
    # ... without a config the caller's (the reference: a random UUIDv4)
    src.write_text(po.to_json_compact(_doc(res, run_id=None)))
    _tab(str(src), str(tmp_path / "b.tsv"), run_id=RUN_ID)
    assert (tmp_path / "b.tsv").read_text() == po.results_to_tabular(res, RUN_ID, to_stdout=False)
    _tab(str(src), str(tmp_path / "c.tsv"))
    import re

    rid = re.search(r"[0-9a-f]{8}-[0-9a-f]{4}-4[0-9a-f]{3}-[89ab][0-9a-f]{3}-[0-9a-f]{12}", (tmp_path / "c.tsv").read_text()).group(0)
    assert (tmp_path / "c.tsv").read_text() == po.results_to_tabular(res, rid, to_stdout=False)
    # upper-case / simple-form UUIDs are printed in Uuid's Display form
    src.write_text(po.to_json_compact(_doc(res, run_id=RUN_ID.upper().replace("-", ""))))
    _tab(str(src), str(tmp_path / "d.tsv"))
    assert (tmp_path / "d.tsv").read_text() == po.results_to_tabular(res, RUN_ID, to_stdout=False)


def test_jsonl_input(tmp_path):
    from blutils_b200 import MappedErrors

    res = _results()
    body = po.results_to_jsonl(res, RUN_ID)
    src = tmp_path / "r.jsonl"
    (tmp_path / "r.json").write_text("{}")  # the reference's existence check looks at <name>.json whatever the format
    # run-with-consensus style: a config line (recognised by `isConfig`), blank lines skipped
    src.write_text(json.dumps({"runId": OTHER_ID, "isConfig": True}) + "\n\n" + body)
    _tab(str(src), str(tmp_path / "r.tsv"), "jsonl")
    assert (tmp_path / "r.tsv").read_text() == po.results_to_tabular(res, RUN_ID, to_stdout=False)
    # build-consensus writes `null` as the config line: not a QueryWithConsensus -> the reference fails, so does this
    src.write_text("null\n" + body)
    with pytest.raises(MappedErrors, match="unable to parse line as JSON"):
        _tab(str(src), str(tmp_path / "r2.tsv"), "jsonl")


def test_errors(tmp_path):
    from blutils_b200 import MappedErrors, Unsupported

    with pytest.raises(MappedErrors, match="does not exist"):
        _tab(str(tmp_path / "absent.json"), None)
    res = _results()
    only_jsonl = tmp_path / "x.jsonl"
    only_jsonl.write_text(po.results_to_jsonl(res, RUN_ID))
    with pytest.raises(MappedErrors, match="does not exist"):  # x.json is what the reference probes
        _tab(str(only_jsonl), None, "jsonl")
    bad = tmp_path / "bad.json"
    for text, what in [("{\"config\": null}", "missing field `results`"), ("{\"results\": [{\"taxon\": null}], \"config\": null}", "missing field `query`"),
                       ("{\"results\": [{\"query\": \"q\", \"query\": \"q\"}]}", "duplicate field `query`"),
                       ("{\"results\": [{\"query\": \"q\", \"taxon\": {\"reachedRank\": \"species\"}}]}", "missing field `identifier`"),
                       ("{\"results\": [{\"query\": \"q\", \"runId\": \"not-a-uuid\"}]}", "invalid UUID"), ("{\"results\": [] } x", "trailing characters"),
                       ("[]", "unexpected character")]:
        bad.write_text(text)
        with pytest.raises(MappedErrors, match=what):
            _tab(str(bad), str(tmp_path / "o.tsv"))
    bad.write_text("{\"results\": []}")
    _tab(str(bad), str(tmp_path / "empty.tsv"))  # config may be missing (Option); no results -> header only
    assert (tmp_path / "empty.tsv").read_text() == po.results_to_tabular([], RUN_ID, to_stdout=False)


def test_extra_fields_floats_and_ranks(tmp_path):
    """Unknown fields are skipped, Option fields may be absent, f64 fields take integers / exponents, rank strings pass through."""
    doc = {"results": [{"query": "q1", "extra": {"a": [1, 2]}, "taxon": {
        "reachedRank": "species-group", "identifier": "x y", "percIdentity": 1e2, "bitScore": 845, "mutated": False, "singleMatch": True,
        "consensusBeans": [{"rank": "Species", "identifier": "b", "occurrences": 3, "accessions": []},
                           {"rank": "genus", "identifier": "c", "occurrences": 1, "taxonomy": "d__x;g__c", "accessions": ["A.1", "B.2"]}]}},
        {"runId": None, "query": "q2", "taxon": {"reachedRank": "genus", "maxAllowedRank": None, "identifier": "g", "percIdentity": 97.125,
                                                  "bitScore": 0.5, "taxonomy": "d__x;g__g", "mutated": True, "singleMatch": False,
                                                  "consensusBeans": None}}]}
    src = tmp_path / "m.json"
    src.write_text(json.dumps(doc))
    _tab(str(src), str(tmp_path / "m.tsv"), run_id=RUN_ID)
    for r in doc["results"]:
        for b in r["taxon"].get("consensusBeans") or []:
            b.setdefault("taxonomy", None)
    want = po.results_to_tabular([{"query": r["query"], "taxon": dict({"taxonomy": None, "consensusBeans": None}, **r["taxon"])} for r in doc["results"]],
                                 RUN_ID, to_stdout=False)
    got = (tmp_path / "m.tsv").read_text()
    assert got == want
    assert "\t100\t845\t" in got and "\tspecies-group\tx y\t" in got and "\tSpecies\tb\tnull\t845\tnull\t" in got and "\t97.125\t0.5\t" in got


def test_cli_stdout_and_stdin(tmp_path):
    """The CLI shim: result on stdin ("-" is the default), TSV on stdout with one println! per piece."""
    cli = os.path.join(ROOT, "blutils_b200", "blu")
    if not os.path.exists(cli):
        pytest.skip("CLI shim not built")
    res = _results()
    text = po.to_json_pretty(_doc(res))
    p = subprocess.run([cli, "blastn", "build-tabular"], input=text.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert p.returncode == 0, p.stderr
    assert p.stdout.decode() == po.results_to_tabular(res, RUN_ID, to_stdout=True)
    src = tmp_path / "in.json"
    src.write_text(text)
    p = subprocess.run([cli, "blastn", "build-tabular", str(src), "-o", str(tmp_path / "out.tsv"), "--input-format=json"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=60)
    assert p.returncode == 0 and p.stdout == b"" and (tmp_path / "out.tsv").read_text() == po.results_to_tabular(res, RUN_ID, to_stdout=False)
    p = subprocess.run([cli, "blastn", "build-tabular", str(tmp_path / "nope.json")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=60)
    assert p.returncode == 101 and b"does not exist" in p.stderr


def _random_results(seed, n=300):
    import random

    rng = random.Random(seed)
    alphabet = ["a", "Z", "0", " ", "_", "-", ".", ";", "\"", "\\", "/", "\t", "\n", "é", "ß", "漢", "𝔘", " ", "|", "'"]

    def word(lo=1, hi=12):
        return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))

    def number():
        k = rng.random()
        if k < 0.3:
            return float(rng.randint(0, 2000))
        if k < 0.6:
            return round(rng.uniform(0, 100), rng.randint(0, 6))
        if k < 0.8:
            return rng.uniform(0, 1e-5)
        return rng.uniform(1e15, 1e22)

    def shuffled(d):
        items = list(d.items())
        rng.shuffle(items)
        return dict(items)

    results = []
    for i in range(n):
        if rng.random() < 0.15:
            results.append({"query": word(), "taxon": None})
            continue
        beans = None
        if rng.random() < 0.7:
            beans = [{"rank": rng.choice(["species", "genus", "clade", "species-group", word()]), "identifier": word(), "occurrences": rng.randint(1, 500),
                      "taxonomy": rng.choice([None, "d__x;p__" + word()]), "accessions": [word() for _ in range(rng.randint(0, 4))]}
                     for _ in range(rng.randint(0, 5))]
        results.append({"query": word(), "taxon": {
            "reachedRank": rng.choice(["domain", "family", "strain", word()]), "maxAllowedRank": rng.choice([None, "genus"]), "identifier": word(),
            "percIdentity": number(), "bitScore": number(), "taxonomy": rng.choice([None, "d__bac;" + word()]), "mutated": rng.random() < 0.5,
            "singleMatch": rng.random() < 0.5, "consensusBeans": beans}})
    return rng, results, shuffled


def test_random_documents(tmp_path):
    """Seeded random result documents (escapes, non-ASCII, surrogate pairs, float shapes, optional fields in any order) written by
    Python's json module in both ASCII-escaped and raw UTF-8 form: same TSV as the oracle computes from the same objects."""
    rng, results, shuffled = _random_results(20261018)
    doc = {"config": None, "results": [shuffled(dict(r, runId=RUN_ID, taxon=shuffled(r["taxon"]) if r["taxon"] else None)) for r in results]}
    want = po.results_to_tabular(results, RUN_ID, to_stdout=False)
    for ascii_only in (True, False):
        src = tmp_path / f"rand{int(ascii_only)}.json"
        src.write_text(json.dumps(doc, ensure_ascii=ascii_only, indent=rng.choice([None, 1, 4])), encoding="utf-8")
        _tab(str(src), str(tmp_path / f"rand{int(ascii_only)}.tsv"))
        assert (tmp_path / f"rand{int(ascii_only)}.tsv").read_text(encoding="utf-8") == want


# ---- YAML input (file_or_stdin.rs:113-130: serde_yaml::from_str::<BlutilsOutput>) ----------------------------------------------

def _yaml_doc(results, run_id=RUN_ID, seq_indent=0, config="null"):
    """The block style serde_yaml writes (pyoracle.results_to_yaml), with Option fields that are None as `null` and, optionally,
    sequences indented below their key (the other common style)."""
    def s(v):
        return "null" if v is None else po.yaml_str(v)

    si = " " * seq_indent
    o = ["results: []\n" if not results else "results:\n"]
    for r in results:
        p = si + "  "
        o.append(si + "- " + (("runId: " + s(run_id) + "\n" + p) if run_id else "") + "query: " + s(r["query"]) + "\n")
        t = r["taxon"]
        if t is None:
            o.append(p + "taxon: null\n")
            continue
        q = p + "  "
        o.append(p + "taxon:\n" + q + "reachedRank: " + s(t["reachedRank"]) + "\n" + q + "maxAllowedRank: " + s(t["maxAllowedRank"]) + "\n" + q +
                 "identifier: " + s(t["identifier"]) + "\n" + q + "percIdentity: " + po.ryu_f64(t["percIdentity"]) + "\n" + q + "bitScore: " +
                 po.ryu_f64(t["bitScore"]) + "\n" + q + "taxonomy: " + s(t["taxonomy"]) + "\n" + q + "mutated: " + ("true" if t["mutated"] else "false") +
                 "\n" + q + "singleMatch: " + ("true" if t["singleMatch"] else "false") + "\n")
        beans = t["consensusBeans"]
        if beans is None:
            o.append(q + "consensusBeans: null\n")
            continue
        o.append(q + "consensusBeans: []\n" if not beans else q + "consensusBeans:\n")
        for b in beans:
            bi = q + si
            o.append(bi + "- rank: " + s(b["rank"]) + "\n" + bi + "  identifier: " + s(b["identifier"]) + "\n" + bi + "  occurrences: " +
                     str(b["occurrences"]) + "\n" + bi + "  taxonomy: " + s(b["taxonomy"]) + "\n")
            o.append(bi + "  accessions: []\n" if not b["accessions"] else bi + "  accessions:\n")
            for a in b["accessions"]:
                o.append(bi + "  " + si + "- " + s(a) + "\n")
    o.append("config: " + config + "\n")
    return "".join(o)


@pytest.mark.parametrize("seq_indent", [0, 2])
def test_yaml_input_mock_run(tmp_path, seq_indent):
    """The YAML that build-consensus writes for the mock 16S run -> the same TSV as from its JSON."""
    res = _results()
    assert _yaml_doc(res) == po.results_to_yaml(res, RUN_ID)  # (the helper really is the writer's format)
    src = tmp_path / "blutils.yaml"
    (tmp_path / "blutils.json").write_text("{}")  # what the reference's existence check looks at (mod.rs:24-33)
    src.write_text(_yaml_doc(res, seq_indent=seq_indent))
    _tab(str(src), str(tmp_path / "t.tsv"), "yaml")
    assert (tmp_path / "t.tsv").read_text() == po.results_to_tabular(res, RUN_ID, to_stdout=False)
    # through the CLI, from stdin
    cli = os.path.join(ROOT, "blutils_b200", "blu")
    if os.path.exists(cli):
        p = subprocess.run([cli, "blastn", "build-tabular", "--input-format", "yaml"], input=src.read_bytes(), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           timeout=60)
        assert p.returncode == 0, p.stderr
        assert p.stdout.decode() == po.results_to_tabular(res, RUN_ID, to_stdout=True)


def test_yaml_input_random_documents(tmp_path):
    """Random strings (quotes, escapes, non-ASCII, astral characters, YAML indicators, things that look like numbers / booleans /
    null) through the writer's quoting rules (pyoracle.yaml_str) and back: the TSV is the one computed from the objects."""
    rng, results, _ = _random_results(777, n=400)
    tricky = ["null", "~", "true", "No", "0x1F", "1e3", "-.inf", "007", "- x", "a: b", "#c", "x #y", "'q'", "\"d\"", " lead", "trail ", "[a]", "{b}", "&a", "*a",
              "!t", "|", ">", "%", "@", "`", "?", ": ", "---", "...", "a\tb", "\u00e9\u2028x", "\U0001d518", "back\\slash"]
    for i, r in enumerate(results):
        r["query"] = tricky[i % len(tricky)] if i % 3 == 0 else r["query"]
        if r["taxon"] and i % 5 == 0:
            r["taxon"]["identifier"] = tricky[(i // 5) % len(tricky)]
    (tmp_path / "r.json").write_text("{}")
    src = tmp_path / "r.yaml"
    src.write_text(_yaml_doc(results), encoding="utf-8")
    _tab(str(src), str(tmp_path / "r.tsv"), "yaml")
    assert (tmp_path / "r.tsv").read_text(encoding="utf-8") == po.results_to_tabular(results, RUN_ID, to_stdout=False)


def test_yaml_input_hand_edited(tmp_path):
    """What a person does to such a file: comments, a document start, blank lines, other quoting, `~`, unknown keys, a config with
    a run id, integers and exponents where floats are expected, deeper indentation."""
    text = """# produced by blu, edited by hand
---
config:
    runId: "11111111-2222-4333-8444-555555555555"   # used where a result has none
    isConfig: true
    anything: []

results:
    - query: 'q one'
      extra: {}
      taxon:
          reachedRank: species-group
          identifier: "x\ty \u00e9 \x41"
          percIdentity: 1e2
          bitScore: 845
          taxonomy: ~
          mutated: False
          singleMatch: TRUE
          consensusBeans:
              - rank: Species
                identifier: b
                occurrences: 0x10
                accessions: [ ]
              - {}
    - runId: 0b0e3c55-7a2f-4a61-9d4e-5f1c2a7b8c9d
      query: q2
      taxon: null
    - query: "3"
      runId: null
      taxon:
          reachedRank: genus
          maxAllowedRank:
          identifier: g
          percIdentity: 97.125
          bitScore: +.5
          mutated: true
          singleMatch: false
...
"""
    (tmp_path / "h.json").write_text("{}")
    src = tmp_path / "h.yaml"
    from blutils_b200 import MappedErrors

    src.write_text(text)
    with pytest.raises(MappedErrors, match="missing field `rank`"):  # the `- {}` bean
        _tab(str(src), str(tmp_path / "h.tsv"), "yaml")
    src.write_text(text.replace("              - {}\n", ""))
    _tab(str(src), str(tmp_path / "h.tsv"), "yaml")
    got = (tmp_path / "h.tsv").read_text().split("\n")  # (to a file the pieces are written without line breaks: one long line)
    want = po.results_to_tabular([
        {"query": "q one", "taxon": {"reachedRank": "species-group", "maxAllowedRank": None, "identifier": "x\ty \u00e9 A", "percIdentity": 100.0, "bitScore": 845.0,
                                     "taxonomy": None, "mutated": False, "singleMatch": True,
                                     "consensusBeans": [{"rank": "Species", "identifier": "b", "occurrences": 16, "taxonomy": None, "accessions": []}]}},
        {"query": "q2", "taxon": None},
        {"query": "3", "taxon": {"reachedRank": "genus", "maxAllowedRank": None, "identifier": "g", "percIdentity": 97.125, "bitScore": 0.5, "taxonomy": None,
                                 "mutated": True, "singleMatch": False, "consensusBeans": None}}], OTHER_ID, to_stdout=False)
    assert "\n".join(got) == want


def test_yaml_input_errors_and_refusals(tmp_path):
    from blutils_b200 import MappedErrors, Unsupported

    (tmp_path / "e.json").write_text("{}")
    src = tmp_path / "e.yaml"
    ok = "results:\n- query: q\n  taxon: null\nconfig: null\n"
    src.write_text(ok)
    _tab(str(src), str(tmp_path / "e.tsv"), "yaml")
    for text, what in [("config: null\n", "missing field `results`"), ("results:\n- taxon: null\n", "missing field `query`"),
                       ("results:\n- query: q\n  query: r\n", "duplicate entry"), ("results: 3\n", "expected a sequence"),
                       ("results:\n- query: q\n  runId: nope\n", "invalid UUID"), ("results:\n- query: [a, b]\n", None),
                       ("results:\n- query: q\n  taxon:\n    reachedRank: s\n    identifier: i\n    percIdentity: high\n", "expected a float"),
                       ("results:\n- query: q\n  taxon:\n    reachedRank: s\n    identifier: i\n    percIdentity: 1\n    bitScore: 2\n    mutated: yes\n",
                        "expected a boolean"),
                       ("results:\n- query: \"a\\qb\"\n", "unknown escape"), ("results:\n\t- query: q\n", "tab"), ("- a\nb: c\n", "document end"),
                       ("results:\n- query: q\n   taxon: null\n", "indentation")]:
        src.write_text(text)
        if what is None:
            with pytest.raises(Unsupported):
                _tab(str(src), str(tmp_path / "e.tsv"), "yaml")
        else:
            with pytest.raises(MappedErrors, match=what):
                _tab(str(src), str(tmp_path / "e.tsv"), "yaml")
    for text in ["results: &a []\n", "results: !!seq []\n", "results:\n- query: |\n    block\n", "results:\n- query: \"two\n    lines\"\n",
                 "results: []\n---\nresults: []\n", "results:\n- query: *q\n", "? complex\n: key\n"]:
        src.write_text(text)
        with pytest.raises(Unsupported):
            _tab(str(src), str(tmp_path / "e.tsv"), "yaml")


def test_yaml_input_property(tmp_path_factory):
    """hypothesis: any strings (YAML indicators, quotes, escapes, non-ASCII, astral, control characters, number / boolean / null
    look-alikes) through the writer's quoting rules and back through the reader: the TSV is the one computed from the objects."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    odd = st.text(alphabet=st.sampled_from(list("abZ09 _-.;:#'\"\\/|&*!%@`?[]{},>~=\t") + ["é", "漢", "\U0001d518", " ", " ", "﻿", "\x7f", "\x01", "\x1b", "\n"]),
                  min_size=0, max_size=12)
    looks = st.sampled_from(["null", "~", "true", "False", "yes", "0x1F", "0o17", "1e3", "-.inf", ".NaN", "007", "12", "-3", "+4", "1.5", "- x", "a: b", "x #y",
                             " lead", "trail ", "---", "...", "? k", ": v", "[a]", "{b}", "&a", "*a", "!t", "|", ">", "%", "@", "`", ""])
    word = st.one_of(odd, looks)
    d = tmp_path_factory.mktemp("y")
    (d / "p.json").write_text("{}")

    @settings(max_examples=int(os.environ.get("BLU_HYP_EXAMPLES", "150")), deadline=None, suppress_health_check=list(HealthCheck),
              derandomize=not os.environ.get("BLU_HYP_RANDOM"), database=None)
    @given(st.data())
    def run(data):
        results = []
        for _ in range(data.draw(st.integers(1, 3))):
            if data.draw(st.integers(0, 4)) == 0:
                results.append({"query": data.draw(word), "taxon": None})
                continue
            beans = None
            if data.draw(st.booleans()):
                beans = [{"rank": data.draw(word), "identifier": data.draw(word), "occurrences": data.draw(st.integers(0, 2 ** 31 - 1)),
                          "taxonomy": data.draw(st.one_of(st.none(), word)), "accessions": data.draw(st.lists(word, max_size=3))}
                         for _ in range(data.draw(st.integers(0, 3)))]
            results.append({"query": data.draw(word), "taxon": {
                "reachedRank": data.draw(word), "maxAllowedRank": data.draw(st.one_of(st.none(), word)), "identifier": data.draw(word),
                "percIdentity": data.draw(st.floats(0, 100, allow_nan=False)), "bitScore": data.draw(st.floats(0, 1e12, allow_nan=False)),
                "taxonomy": data.draw(st.one_of(st.none(), word)), "mutated": data.draw(st.booleans()), "singleMatch": data.draw(st.booleans()),
                "consensusBeans": beans}})
        src = d / "p.yaml"
        src.write_text(_yaml_doc(results, seq_indent=data.draw(st.sampled_from([0, 2]))), encoding="utf-8")
        _tab(str(src), str(d / "p.tsv"), "yaml")
        assert (d / "p.tsv").read_text(encoding="utf-8") == po.results_to_tabular(results, RUN_ID, to_stdout=False)

    run()


def test_json_input_property(tmp_path_factory):
    """hypothesis: result documents with any strings (every escape json.dumps produces, raw UTF-8, astral and control characters),
    unknown fields holding arbitrary JSON, keys in any order, JSON and JSONL framing: the TSV is the one computed from the objects."""
    import random as _random

    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    text_s = st.text(alphabet=st.one_of(st.sampled_from(list("ab;_ -\"\\/\b\f\n\r\t{}[]:,") + ["é", "漢", "\U0001d518", "\x00", "\x1f", "\x7f", "퟿", "￿"]),
                                        st.characters(blacklist_categories=("Cs",))), max_size=10)
    junk = st.recursive(st.one_of(st.none(), st.booleans(), st.integers(-10 ** 20, 10 ** 20), st.floats(allow_nan=False, allow_infinity=False), text_s),
                        lambda c: st.one_of(st.lists(c, max_size=3), st.dictionaries(text_s, c, max_size=3)), max_leaves=6)
    d = tmp_path_factory.mktemp("jp")

    @settings(max_examples=int(os.environ.get("BLU_HYP_EXAMPLES", "150")), deadline=None, suppress_health_check=list(HealthCheck),
              derandomize=not os.environ.get("BLU_HYP_RANDOM"), database=None)
    @given(st.data())
    def run(data):
        rnd = _random.Random(data.draw(st.integers(0, 2 ** 32)))

        def shuffled(x, extra=True):
            items = list(x.items())
            if extra and rnd.random() < 0.3:
                items.append(("zzUnknown", data.draw(junk)))
            rnd.shuffle(items)
            return dict(items)

        results = []
        for _ in range(data.draw(st.integers(0, 3))):
            if data.draw(st.integers(0, 4)) == 0:
                results.append({"query": data.draw(text_s), "taxon": None})
                continue
            beans = None
            if data.draw(st.booleans()):
                beans = [{"rank": data.draw(text_s), "identifier": data.draw(text_s), "occurrences": data.draw(st.integers(-2 ** 31, 2 ** 31 - 1)),
                          "taxonomy": data.draw(st.one_of(st.none(), text_s)), "accessions": data.draw(st.lists(text_s, max_size=3))}
                         for _ in range(data.draw(st.integers(0, 3)))]
            results.append({"query": data.draw(text_s), "taxon": {
                "reachedRank": data.draw(text_s), "maxAllowedRank": data.draw(st.one_of(st.none(), text_s)), "identifier": data.draw(text_s),
                "percIdentity": data.draw(st.floats(0, 100, allow_nan=False)), "bitScore": data.draw(st.one_of(st.floats(0, 1e15, allow_nan=False), st.integers(0, 10 ** 6))),
                "taxonomy": data.draw(st.one_of(st.none(), text_s)), "mutated": data.draw(st.booleans()), "singleMatch": data.draw(st.booleans()),
                "consensusBeans": beans}})
        want = po.results_to_tabular([dict(r, taxon=dict(r["taxon"], bitScore=float(r["taxon"]["bitScore"])) if r["taxon"] else None) for r in results],
                                     RUN_ID, to_stdout=False)
        objs = [shuffled(dict(r, runId=RUN_ID, taxon=(shuffled(dict(r["taxon"], consensusBeans=[shuffled(b) for b in r["taxon"]["consensusBeans"]]
                                                                                  if r["taxon"]["consensusBeans"] is not None else None)) if r["taxon"] else None)))
                for r in results]
        ascii_only = data.draw(st.booleans())
        src = d / "r.json"
        src.write_text(json.dumps(shuffled({"results": objs, "config": None}), ensure_ascii=ascii_only, indent=data.draw(st.sampled_from([None, 1, 3]))), encoding="utf-8")
        _tab(str(src), str(d / "r.tsv"))
        assert (d / "r.tsv").read_bytes().decode("utf-8") == want
        # the same objects as JSONL (one compact object per line; strings with a raw newline cannot occur: json.dumps escapes them)
        (d / "l.json").write_text("{}")
        srcl = d / "l.jsonl"
        srcl.write_text("".join(json.dumps(o, ensure_ascii=ascii_only) + "\n" for o in objs), encoding="utf-8")
        if not any("isConfig" in json.dumps(o, ensure_ascii=ascii_only) for o in objs):  # (a line holding that word is taken for the config line)
            _tab(str(srcl), str(d / "l.tsv"), "jsonl")
            assert (d / "l.tsv").read_bytes().decode("utf-8") == want

    run()
