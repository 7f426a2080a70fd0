"""Binary side-car cache of the encoded lineage tables.  The reference re-parses the taxonomy JSON on
every run (build_consensus_identities/mod.rs:254-265); the cache must give exactly the tables a fresh parse gives, and must
never be used when the JSON, `use_taxid` or the cutoffs differ, or when the cache file is damaged."""
import json
import os

import pytest

import sim_ffi

HERE = os.path.dirname(os.path.abspath(__file__))
MOCK = os.path.join(HERE, "golden", "mock16s", "mock-16S.blutils.json")


def _small_json(path, n=200, tweak=None):
    taxa = []
    for i in range(n):
        lin = f"d__bac;p__p{i % 3};c__c{i % 7};o__o{i % 11};f__f{i % 13};g__g{i % 50};s__s{i}"
        if i % 9 == 0:
            lin = lin.replace(";p__", ";clade__x%d;p__" % (i % 4))
        if i % 17 == 0:
            lin += ";strain__st%d" % i
        taxa.append({"taxid": 1000 + i * 7, "rank": "species", "numericLineage": "d__2;s__%d" % (1000 + i * 7), "textLineage": lin, "accessions": []})
    if tweak:
        tweak(taxa)
    with open(path, "w") as f:
        json.dump({"blutilsVersion": "8.3.1", "ignoreTaxids": None, "replaceRank": None, "dropNonLinnaeanTaxonomies": None,
                   "sourceDatabase": "test", "taxonomies": taxa}, f)
    return path


def test_cache_round_trip(tmp_path):
    js = _small_json(str(tmp_path / "t.json"))
    cache = str(tmp_path / "t.json.blucache")
    s0, ck0 = sim_ffi.taxonomy_cached(js, cache)
    assert s0 == 0 and os.path.exists(cache)
    s1, ck1 = sim_ffi.taxonomy_cached(js, cache)
    assert s1 == 1 and ck1 == ck0
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]  # the temporary file was renamed into place


@pytest.mark.skipif(not os.path.exists(MOCK), reason="mock 16S fixture missing")
def test_cache_round_trip_mock16s(tmp_path):
    cache = str(tmp_path / "mock.blucache")
    for use_taxid in (False, True):
        c = cache + str(int(use_taxid))
        s0, ck0 = sim_ffi.taxonomy_cached(MOCK, c, use_taxid=use_taxid)
        s1, ck1 = sim_ffi.taxonomy_cached(MOCK, c, use_taxid=use_taxid)
        assert (s0, s1) == (0, 1) and ck0 == ck1


def test_cache_key_covers_everything_the_tables_depend_on(tmp_path):
    js = _small_json(str(tmp_path / "t.json"))
    cache = str(tmp_path / "c.bin")
    _, ck_bac = sim_ffi.taxonomy_cached(js, cache, taxon="bacteria")
    # other cutoff backbone: the interpolated cutoffs differ, so the cache must not be reused
    s, ck_fun = sim_ffi.taxonomy_cached(js, cache, taxon="fungi")
    assert s == 0 and ck_fun != ck_bac
    s, ck = sim_ffi.taxonomy_cached(js, cache, taxon="fungi")
    assert s == 1 and ck == ck_fun
    # custom values
    cust = {"domain": 60, "kingdom": None, "phylum": 70, "class": 75, "order": 80, "family": 85, "genus": 90, "species": 97}
    s, ck_c1 = sim_ffi.taxonomy_cached(js, cache, taxon="custom", custom=cust)
    assert s == 0
    s, ck = sim_ffi.taxonomy_cached(js, cache, taxon="custom", custom=dict(cust, species=98))
    assert s == 0 and ck != ck_c1
    # use_taxid picks the other lineage column
    s, ck_num = sim_ffi.taxonomy_cached(js, cache, taxon="bacteria", use_taxid=True)
    assert s == 0 and ck_num != ck_bac
    # same size, different content of the JSON
    _small_json(js, tweak=lambda taxa: taxa[5].update(textLineage=taxa[5]["textLineage"].replace("g__g5", "g__g6")))
    s, ck_mod = sim_ffi.taxonomy_cached(js, cache, taxon="bacteria", use_taxid=False)
    assert s == 0 and ck_mod != ck_bac


def test_damaged_cache_is_rebuilt(tmp_path):
    js = _small_json(str(tmp_path / "t.json"))
    cache = str(tmp_path / "c.bin")
    _, ck0 = sim_ffi.taxonomy_cached(js, cache)
    good = open(cache, "rb").read()
    # truncated
    open(cache, "wb").write(good[: len(good) // 2])
    s, ck = sim_ffi.taxonomy_cached(js, cache)
    assert s == 0 and ck == ck0 and open(cache, "rb").read() == good
    # one flipped payload byte
    bad = bytearray(good)
    bad[len(bad) * 3 // 4] ^= 0x40
    open(cache, "wb").write(bytes(bad))
    s, ck = sim_ffi.taxonomy_cached(js, cache)
    assert s == 0 and ck == ck0
    # trailing garbage
    open(cache, "wb").write(good + b"xx")
    s, ck = sim_ffi.taxonomy_cached(js, cache)
    assert s == 0 and ck == ck0
    # not a cache at all / empty
    for junk in (b"", b"hello world, definitely not a cache file" * 10):
        open(cache, "wb").write(junk)
        s, ck = sim_ffi.taxonomy_cached(js, cache)
        assert s == 0 and ck == ck0
    s, ck = sim_ffi.taxonomy_cached(js, cache)
    assert s == 1 and ck == ck0


def test_unwritable_cache_location_still_loads(tmp_path):
    js = _small_json(str(tmp_path / "t.json"))
    s, ck = sim_ffi.taxonomy_cached(js, str(tmp_path / "no_such_dir" / "c.bin"))
    s2, ck2 = sim_ffi.taxonomy_cached(js, str(tmp_path / "c.bin"))
    assert s == -1 and s2 == 0 and ck == ck2


def test_missing_json_is_an_io_error(tmp_path):
    with pytest.raises(IOError):
        sim_ffi.taxonomy_cached(str(tmp_path / "absent.json"), str(tmp_path / "c.bin"))
    # even with a valid-looking cache present: the key needs the JSON's hash
    js = _small_json(str(tmp_path / "t.json"))
    sim_ffi.taxonomy_cached(js, str(tmp_path / "c.bin"))
    os.remove(js)
    with pytest.raises(IOError):
        sim_ffi.taxonomy_cached(js, str(tmp_path / "c.bin"))


def test_empty_taxonomy_list(tmp_path):
    js = str(tmp_path / "e.json")
    with open(js, "w") as f:
        json.dump({"blutilsVersion": "8.3.1", "sourceDatabase": "x", "taxonomies": []}, f)
    s0, ck0 = sim_ffi.taxonomy_cached(js, str(tmp_path / "e.bin"))
    s1, ck1 = sim_ffi.taxonomy_cached(js, str(tmp_path / "e.bin"))
    assert (s0, s1) == (0, 1) and ck0 == ck1


@pytest.mark.gpu
@pytest.mark.parametrize("use_taxid,strategy", [(False, "relaxed"), (True, "cautious")])
def test_engine_loaded_from_cache_gives_the_reference_results(tmp_path, use_taxid, strategy):
    """C ABI blu_taxonomy_load_json_cached: the mock 16S run through tables built + cached, then through tables read back
    from the cache, against the committed expected output."""
    from blutils_b200 import ConsensusEngine, ConsensusStrategy, Taxon

    d = os.path.join(HERE, "golden", "mock16s")
    want = open(os.path.join(d, f"expected.{'taxid' if use_taxid else 'text'}.{strategy}.jsonl"), "rb").read()
    headers = open(os.path.join(d, "headers.txt")).read().split()
    cache = str(tmp_path / "mock.blucache")
    states = []
    for _ in range(2):
        eng = ConsensusEngine(Taxon.Bacteria, ConsensusStrategy.Cautious if strategy == "cautious" else ConsensusStrategy.Relaxed, use_taxid, None)
        states.append(eng.load_taxonomy(MOCK, cache=cache))
        out = eng.run_file(os.path.join(d, "blast.out"))
        out.add_headers(headers)
        assert out.jsonl() == want
        out.close()
        eng.close()
    assert states == [0, 1]
    # a context with other cutoffs must not accept this cache
    eng = ConsensusEngine(Taxon.Fungi, ConsensusStrategy.Relaxed, use_taxid, None)
    assert eng.load_taxonomy(MOCK, cache=cache) == 0
    eng.close()
