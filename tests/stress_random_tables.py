#!/usr/bin/env python
"""Randomised GPU stress of the streaming path against the C++ oracle: many mid-sized tables (a few MB each: hundreds of 32 KB
windows, several CTA segments) with random hit counts, tie rates, id lengths, number shapes, duplicate rows, with and without a
final newline, contiguous and scattered; each through the host-text path (default chunks and 1 MiB chunks: the carry between
chunks), the device-text path and the device-resident path.  Not collected by pytest and not part of bench.py; it lives under
tests/ because it uses the oracle as its checker.

  python tests/stress_random_tables.py --tables 40 --seed 1
"""
import argparse, json, os, random, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tables", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    import torch
    from blutils_b200 import ConsensusEngine, ConsensusPanic, ConsensusStrategy, Taxon
    from helpers import random_blast, random_taxonomy
    from oracle_ffi import Oracle, OracleDataError

    rng = random.Random(a.seed)
    n_ok = n_abort = 0
    rows_total = bytes_total = 0
    t0 = time.time()
    for k in range(a.tables):
        units = random_taxonomy(rng, n_leaves=rng.choice([8, 40, 200]), shared_root=True)
        ids = [u["taxid"] for u in units]
        lin = [u["textLineage"] for u in units]
        text = random_blast(rng, units, n_queries=rng.choice([300, 2000, 6000]), max_hits=rng.choice([1, 3, 12, 40, 150]), contiguous=rng.random() < 0.85,
                            tie_rate=rng.choice([0.1, 0.6, 0.95]), low_pident=rng.choice([60.0] * 9 + [45.0]))
        pad = rng.choice([0, 0, 7, 33, 120])
        if pad:  # longer ids: other row lengths, other window alignments
            text = text.replace(b"draft-", b"draft-" + b"x" * pad).replace(b"SRR1.", b"SRR1." + b"y" * (pad // 2))
        strategy = rng.choice(["cautious", "relaxed"])
        try:
            want = Oracle(ids, lin, "bacteria", strategy, None, threads=os.cpu_count()).run_raw(text)[0]
        except OracleDataError:
            want = None
        rows_total += text.count(b"\n")
        bytes_total += len(text)
        strat = ConsensusStrategy.Cautious if strategy == "cautious" else ConsensusStrategy.Relaxed
        for chunk in (0, 1 << 20):
            eng = ConsensusEngine(Taxon.Bacteria, strat, False, None, chunk_bytes=chunk)
            eng.load_taxonomy_arrays(ids, lin)
            t = torch.zeros((len(text) + 255) // 128 * 128, dtype=torch.uint8, device="cuda")
            t[:len(text)] = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
            torch.cuda.synchronize()
            stream = torch.cuda.current_stream().cuda_stream
            calls = [("host", lambda: eng.run_host(text)), ("device", lambda: eng.run_device(t.data_ptr(), len(text), stream))]
            if chunk == 0:
                calls.append(("resident", lambda: eng.run_device_resident(t.data_ptr(), len(text), stream).download()))
            for name, call in calls:
                try:
                    got = call().jsonl()
                except ConsensusPanic:
                    got = None
                if got != want:
                    open("gpurun_out/stress_failure.blast.out", "wb").write(text)
                    json.dump({"ids": ids, "lin": lin, "strategy": strategy, "path": name, "chunk": chunk, "table": k, "seed": a.seed},
                              open("gpurun_out/stress_failure.json", "w"))
                    print(json.dumps({"ok": False, "table": k, "path": name, "chunk": chunk, "want_abort": want is None, "got_abort": got is None}))
                    sys.exit(3)
            eng.close()
        if want is None:
            n_abort += 1
        else:
            n_ok += 1
    print(json.dumps({"ok": True, "seed": a.seed, "tables": a.tables, "with_result": n_ok, "reference_aborts": n_abort, "rows": rows_total,
                      "text_mb": round(bytes_total / 1e6, 1), "paths": "host (default and 1 MiB chunks), device text (both), device-resident + download",
                      "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
