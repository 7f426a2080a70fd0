"""Query-aligned shard cuts: the in-memory and the file variant agree, never split a query, cover the table, and
cope with queries longer than the read window, empty lines, a missing last newline and more shards than queries."""
import random

import pytest

from blutils_b200 import MappedErrors, shard_cuts, shard_cuts_file


def _table(rng, n_queries, max_rows, long_query=None):
    rows = []
    for q in range(n_queries):
        n = long_query[1] if long_query and q == long_query[0] else rng.randint(1, max_rows)
        qid = f"q{q:05d}" + ("x" * rng.randint(0, 30))
        for h in range(n):
            rows.append(f"{qid}\tNR_{h}.1\t{rng.randint(1, 99999)}\t99.{h % 10}\t400\t0\t0\t1\t400\t1\t400\t0.0\t{700 - h}")
            if rng.random() < 0.02:
                rows.append("")  # an empty line
    return rows


@pytest.mark.parametrize("seed", range(6))
def test_cuts_memory_equals_file(tmp_path, seed):
    rng = random.Random(seed)
    long_q = (rng.randint(0, 199), 40_000) if seed % 2 else None  # ~2.5 MB query: longer than the file variant's 1 MB window
    rows = _table(rng, 200, 60, long_q)
    text = ("\n".join(rows) + ("\n" if seed % 3 else "")).encode()
    path = tmp_path / "t.out"
    path.write_bytes(text)
    starts = {0}
    pos = 0
    prev = None
    for line in text.split(b"\n"):
        if line:
            q = line.split(b"\t")[0]
            if q != prev:
                starts.add(pos)
            prev = q
        pos += len(line) + 1
    for n in (1, 2, 3, 8, 64, 500):
        a = shard_cuts(text, n)
        b = shard_cuts_file(str(path), n)
        assert a == b, n
        assert a[0] == 0 and a[-1] == len(text) and a == sorted(a)
        for c in a[1:-1]:
            assert c == len(text) or c in starts, (n, c)  # every cut is the first byte of a query's first row


def test_cuts_file_errors(tmp_path):
    with pytest.raises(MappedErrors):
        shard_cuts_file(str(tmp_path / "missing.out"), 4)
    p = tmp_path / "empty.out"
    p.write_bytes(b"")
    assert shard_cuts_file(str(p), 3) == [0, 0, 0, 0]
