"""Builds BASELINE config C1: the reference's mock 16S database + a hand-authored outfmt-6 hit table.

Input (reference repo, read-only, not present on the GPU box -> run here and commit the outputs):
    /root/reference/test/mock/input/ref_databases/mock-16S_taxonomies.tsv   legacy `accession<TAB>numeric lineage`
    /root/reference/test/mock/input/query/query.fna                         the 10 mock query ids
The legacy TSV (join by accession) is not readable by blutils 8.3.1, whose taxonomy file is the `TaxonomiesMap`
JSON joined by integer taxid (core/src/domain/dtos/taxonomies_map.rs:6-32).  Conversion: one synthetic taxid per
distinct lineage string (900001, 900002, ... in order of first appearance), accessions grouped under it,
`numericLineage` = the TSV lineage, `textLineage` = the same lineage with `n<id>` identifiers.

`blastn` is not available, so the hit table is authored by hand (function `hits()`): it covers, per SURVEY 8c(3),
single hit; all-agree to species; divergence at species / genus / order / phylum; lineages of unequal length under both
strategies; `strain`, `species-group`, `species-subgroup`, `clade` ranks; a query without hits (INVALID_SEQUENCE, only
in `headers`); fractional bit scores falling into one truncated group (84.2 / 84.9); ties on every sort key;
duplicate accession rows (dedup).  Expected outputs are produced by oracle/pyoracle.py and cross-checked against
oracle/blu_oracle.cpp.

    python tests/golden/make_mock16s.py
"""
import collections
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
OUT = os.path.join(HERE, "mock16s")
TSV = "/root/reference/test/mock/input/ref_databases/mock-16S_taxonomies.tsv"
FNA = "/root/reference/test/mock/input/query/query.fna"


def main():
    import pyoracle as po
    from oracle_ffi import Oracle

    os.makedirs(OUT, exist_ok=True)
    by_lineage = collections.OrderedDict()
    for line in open(TSV):
        line = line.rstrip("\n")
        if not line:
            continue
        acc, lin = line.split("\t")
        by_lineage.setdefault(lin, [])
        if acc not in by_lineage[lin]:
            by_lineage[lin].append(acc)
    units, taxid_of = [], {}
    for i, (lin, accs) in enumerate(by_lineage.items()):
        taxid = 900001 + i
        taxid_of[lin] = taxid
        text = ";".join(f"{p.split('__')[0]}__n{p.split('__')[1]}" for p in lin.split(";"))
        units.append({"taxid": taxid, "rank": lin.split(";")[-1].split("__")[0], "numericLineage": lin, "textLineage": text,
                      "accessions": [{"accession": a, "oid": str(k)} for k, a in enumerate(accs)]})
    db = {"blutilsVersion": "8.3.1", "ignoreTaxids": None, "replaceRank": None, "dropNonLinnaeanTaxonomies": False,
          "sourceDatabase": "test/mock/input/ref_databases/mock-16S", "taxonomies": units}
    json.dump(db, open(os.path.join(OUT, "mock-16S.blutils.json"), "w"), indent=1)
    queries = [l[1:].strip() for l in open(FNA) if l.startswith(">")]
    assert len(queries) == 10

    def T(lin):
        return taxid_of[lin]

    P = "d__2;p__1224;c__1236;o__135622;f__267890;g__22;"          # Shewanella-like genus 22
    A = "d__2;p__201174;c__1760;o__85006;f__1268;g__1742989;"      # genus 1742989
    B = "d__2;clade__1783272;p__1239;c__91061;o__1385;f__186817;g__1386;"
    PS = "d__2;p__1224;c__1236;o__72274;f__135621;g__286;"

    def row(q, acc, lin, pident, length, bits, ev="0.0"):
        return "\t".join([q, acc, str(T(lin)), pident, str(length), "3", "0", "1", str(length), "10", str(9 + length), ev, bits])

    rows = []
    # 1. NR114924.257984.Bac: single top hit (species reached only if pident >= 99)
    q = queries[0]
    rows += [row(q, "NR114924.257984.Baca", A + "s__257984", "99.356", 466, "845"), row(q, "NR_X1.1", A + "s__256701", "97.000", 466, "790"),
             row(q, "NR_X2.1", A + "s__37930", "95.100", 460, "700", "1e-180")]
    # 2. NR025123.135626.Bac: all top hits agree to species (same taxon, duplicate accession rows -> dedup); unequal lengths below
    q = queries[1]
    rows += [row(q, "NR025123.135626.Baca", P + "s__135626", "100.000", 455, "841"), row(q, "NR025123.135626.Baca", P + "s__135626", "100.000", 455, "841"),
             row(q, "NR025123.135626.Bacb", "d__2;p__1224;c__1236;o__135622;f__267890", "100.000", 455, "841"),
             row(q, "NR_LOW.1", P + "s__93973", "91.000", 455, "600", "2.51e-117")]
    # 3. draft-5123: divergence at species inside genus 22 (three species, ties on pident and length -> accession order)
    q = queries[3]
    rows += [row(q, "NR025012.93973.Bac", P + "s__93973", "98.927", 466, "833"), row(q, "NR025443.150120.Bac", P + "s__150120", "98.927", 466, "833"),
             row(q, "NR_A.1", P + "s__93973", "98.927", 466, "833"), row(q, "NR_B.1", P + "s__640633", "98.500", 466, "833"),
             row(q, "NR_C.1", P + "s__640633", "97.000", 300, "500")]
    # 4. close-to-NR_040877: species-group / species-subgroup lineages, divergence below the species group
    q = queries[4]
    rows += [row(q, "NR_115063.1", B + "species-group__653685;species-subgroup__653388;s__260554", "99.356", 466, "845"),
             row(q, "NR_115282.1", B + "species-group__653685;species-subgroup__653388;s__260554", "99.356", 466, "845"),
             row(q, "NR_024693.1", B + "species-group__653685;s__1423", "99.356", 466, "845"),
             row(q, "NR_112725.1", B + "species-group__653685;s__1423", "99.142", 466, "845"),
             row(q, "NR_OTHER.1", B + "species-group__86661;s__2026194", "96.000", 466, "700")]
    # 5. NR_113097.873513: strain-level lineages with two clades in front; fractional bit scores in one truncated group
    q = queries[5]
    S1 = "d__2;clade__1783270;clade__68336;p__976;c__200643;o__171549;f__171552;g__2974251;s__28126;strain__873513"
    S2 = "d__2;clade__1783270;clade__68336;p__976;c__200643;o__171549;f__171552;g__2974251;s__165179;strain__537011"
    rows += [row(q, "NR_113097.1", S1, "100.000", 45, "84.2", "3e-20"), row(q, "NR_113098.1", S2, "100.000", 45, "84.9", "3e-20"),
             row(q, "NR_113099.1", S1, "97.778", 45, "83.99", "1e-19")]
    # 6. draft-8923: divergence at order (o__85006 vs the truncated o__85005 lineage) -> class-level consensus
    q = queries[6]
    rows += [row(q, "NR_D1.1", A + "s__225894", "93.500", 400, "560"), row(q, "NR_D2.1", "d__2;p__201174;c__1760;o__85005", "93.500", 400, "560"),
             row(q, "NR_D3.1", A + "s__37929", "93.100", 400, "560")]
    # 7. draft-1605: divergence at phylum (Proteobacteria vs Actinobacteria) with a clade-bearing lineage as well
    q = queries[7]
    rows += [row(q, "NR_E1.1", P + "s__238836", "88.000", 380, "410"), row(q, "NR_E2.1", A + "s__1522174", "88.250", 380, "410"),
             row(q, "NR_E3.1", B + "s__2880966", "87.900", 380, "410")]
    # 8. draft-893: species-group siblings (Pseudomonas-like genus 286): strain vs species-group lineages of unequal length
    q = queries[8]
    rows += [row(q, "NR_F1.1", PS + "s__312306;strain__384676", "99.800", 470, "860"), row(q, "NR_F2.1", PS + "species-group__136845;s__70775", "99.800", 470, "860"),
             row(q, "NR_F3.1", PS + "s__485895", "99.800", 470, "860"), row(q, "NR_F4.1", PS + "s__312306;strain__384676", "99.800", 471, "860")]
    # 9. draft-2582: one hit far below the genus cutoff (reaches class only)
    q = queries[9]
    rows += [row(q, "NR_G1.1", B + "species-group__86661;s__2338372", "81.250", 300, "250", "4e-60")]
    # INVALID_SEQUENCE (queries[2]) has no hits: it only appears in `headers`
    text = "\n".join(rows) + "\n"
    open(os.path.join(OUT, "blast.out"), "w").write(text)
    open(os.path.join(OUT, "headers.txt"), "w").write("\n".join(queries) + "\n")
    for use_taxid in (True, False):
        tax = po.load_taxonomy(os.path.join(OUT, "mock-16S.blutils.json"), use_taxid)
        for strategy in ("cautious", "relaxed"):
            res = po.build_consensus_identities(text.encode(), tax, "bacteria", strategy, headers=queries)
            js = po.results_to_jsonl(res)
            ids = [u["taxid"] for u in units]
            lin = [u["numericLineage"] if use_taxid else u["textLineage"] for u in units]
            other = Oracle(ids, lin, "bacteria", strategy).run_raw(text.encode(), headers=queries)[0].decode()
            assert js == other, "the two oracles disagree on the mock fixture"
            name = f"expected.{'taxid' if use_taxid else 'text'}.{strategy}.jsonl"
            open(os.path.join(OUT, name), "w").write(js)
            print(name, len(res), "results;", sum(r["taxon"] is None for r in res), "without taxon")


if __name__ == "__main__":
    main()
