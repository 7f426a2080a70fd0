#!/usr/bin/env python
"""Throughput + parity of the larger BASELINE configs (C3, C4) on one GPU, at a size that fits a short GPU session.
Not collected by pytest and not part of bench.py; it lives under tests/ because it uses the CPU oracle as its checker.
Results are quoted in DESIGN.md.

  python tests/run_configs.py C3 --queries 2000000      # 100 hits/query, 2 M-taxon lineage map, bacteria cutoffs
  python tests/run_configs.py C4 --queries 400000       # Zipf(1.1) hits on [1, 5000]

Parity at size: (1) the first `--check` queries are re-run alone and compared with the CPU oracle (checksum of the
canonical JSONL); (2) shard-sum invariance on the full table: the checksums of 8 query-aligned shards add up to the
checksum of the whole."""
import argparse, ctypes as C, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["C3", "C4"])
    ap.add_argument("--queries", type=int, default=0)
    ap.add_argument("--check", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from blutils_b200 import ConsensusEngine, ConsensusStrategy, Taxon, _ffi, shard_cuts
    from blutils_b200.synth import SynthWorkload
    from oracle_ffi import Oracle, checksum_jsonl

    zipf = a.config == "C4"
    hits = 5000 if zipf else 100
    nq = a.queries or (400_000 if zipf else 2_000_000)
    t0 = time.time()
    w = SynthWorkload(2_000_000, seed=20261018 + (4 if zipf else 3))
    ids, off, blob = w.lineages(False)
    eng = ConsensusEngine(Taxon.Bacteria, ConsensusStrategy.Relaxed, False, None)
    eng.load_taxonomy_raw(ids.ctypes.data, off.ctypes.data, blob.ctypes.data, len(ids))
    t_tax = time.time() - t0
    cap = int(nq * (420 if zipf else 100) * 80 * (1.3 if zipf else 1.0)) + (1 << 20)
    pinned = _ffi.lib().blu_host_alloc(cap)
    assert pinned, "pinned allocation failed"
    t0 = time.time()
    nbytes, nrows = w.hits_into(pinned, cap, 0, nq, hits, zipf=zipf)
    t_gen = time.time() - t0
    # parity 1: prefix vs oracle
    lin = [bytes(blob[int(off[i]):int(off[i + 1])]).decode() for i in range(len(ids))]
    orc = Oracle(ids.tolist(), lin, "bacteria", "relaxed")
    pre = w.hits(0, a.check, hits, zipf=zipf)
    want, _, _ = orc.run_raw(pre)
    ok_prefix = eng.run_host(pre).checksum() == checksum_jsonl(want)
    # e2e
    eng.run_host(pinned, nbytes).close()
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter(); out = eng.run_host(pinned, nbytes); ts.append(time.perf_counter() - t0)
        whole = out.checksum() if _ == 0 else whole
        n_out = len(out); tm = eng.timings(); out.close()
    e2e = min(ts)
    # device resident
    dbuf = torch.empty((nbytes + 255) // 128 * 128, dtype=torch.uint8, device="cuda")
    dbuf[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(pinned), dtype=torch.uint8))
    torch.cuda.synchronize()
    eng.run_device(dbuf.data_ptr(), nbytes).close()
    td = []
    for _ in range(a.steps):
        t0 = time.perf_counter(); out = eng.run_device(dbuf.data_ptr(), nbytes); td.append(time.perf_counter() - t0)
        tmd = eng.timings(); same = out.checksum() == whole if _ == 0 else same; out.close()
    dev = min(td)
    # parity 2: shard-sum invariance
    cuts = shard_cuts(pinned, 8, nbytes)
    tot = 0
    for x, y in zip(cuts[:-1], cuts[1:]):
        if y > x:
            p = eng.run_host(pinned + x, y - x); tot = (tot + p.checksum()) % (1 << 64); p.close()
    print(json.dumps({"config": a.config, "queries": nq, "rows": nrows, "text_gb": nbytes / 1e9, "taxa": 2_000_000,
                      "taxonomy_build_s": round(t_tax, 2), "generate_s": round(t_gen, 2),
                      "device_resident": {"queries_per_s": n_out / dev, "rows_per_s": nrows / dev, "ms": dev * 1e3,
                                          "ms_tile_kernel": tmd["ms_tile_kernel"], "ms_longrun_kernel": tmd["ms_longrun_kernel"],
                                          "ms_consensus_gather": tmd["ms_gather_kernel"], "deferred_runs": tmd["n_deferred_runs"],
                                          "tile_kernel_gb_per_s": nbytes / tmd["ms_tile_kernel"] / 1e6},
                      "e2e": {"queries_per_s": n_out / e2e, "rows_per_s": nrows / e2e, "ms": e2e * 1e3, "text_gb_per_s": nbytes / e2e / 1e9},
                      "parity": {"prefix_vs_oracle": ok_prefix, "prefix_queries": a.check, "device_equals_host": same,
                                 "shard_sum_equals_whole": tot == whole}}))


if __name__ == "__main__":
    main()
