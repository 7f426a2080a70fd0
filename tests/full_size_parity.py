#!/usr/bin/env python
"""FULL-output parity at full size, one GPU: every query of the table, not a prefix.  The CPU oracle (16 threads) walks the
same text; the checksum of its canonical JSONL must equal the checksum of the GPU result -- for the host-text path, for the
device-resident path and for the device-resident-then-downloaded path.  Not collected by pytest and not part of bench.py; it
lives under tests/ because it uses the oracle as its checker.  Results are quoted in DESIGN.md section 7.

  python tests/full_size_parity.py c2            # BASELINE configs[1] at its stated size: 1 M queries x 50 hits, 3.8 GB
  python tests/full_size_parity.py zipf          # 60 k queries, Zipf(1.1) hits on [1, 5000], 2 M-taxon map (C4's shape)
  python tests/full_size_parity.py c3 --queries 300000   # 100 hits/query, 2 M-taxon map, cautious
"""
import argparse, ctypes as C, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))

SHAPES = {"c2": dict(queries=1_000_000, hits=50, zipf=False, taxa=30_000, strategy="relaxed", seed=2),
          "c3": dict(queries=300_000, hits=100, zipf=False, taxa=2_000_000, strategy="cautious", seed=3),
          "zipf": dict(queries=60_000, hits=5000, zipf=True, taxa=2_000_000, strategy="relaxed", seed=4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", choices=sorted(SHAPES))
    ap.add_argument("--queries", type=int, default=0)
    a = ap.parse_args()
    import torch
    from blutils_b200 import ConsensusEngine, ConsensusStrategy, Taxon
    from blutils_b200.synth import SynthWorkload
    from oracle_ffi import Oracle, checksum_jsonl

    sh = SHAPES[a.shape]
    nq = a.queries or sh["queries"]
    w = SynthWorkload(sh["taxa"], seed=20261018 + sh["seed"])
    ids, off, blob = w.lineages(False)
    strategy = ConsensusStrategy.Relaxed if sh["strategy"] == "relaxed" else ConsensusStrategy.Cautious
    eng = ConsensusEngine(Taxon.Bacteria, strategy, False, None)
    eng.load_taxonomy_raw(ids.ctypes.data, off.ctypes.data, blob.ctypes.data, len(ids))
    text = w.hits(0, nq, sh["hits"], zipf=sh["zipf"])
    lin = [bytes(blob[int(off[i]):int(off[i + 1])]).decode() for i in range(len(ids))]
    t0 = time.time()
    want, n_q, n_rows = Oracle(ids.tolist(), lin, "bacteria", sh["strategy"], None, threads=os.cpu_count()).run_raw(text)
    t_oracle = time.time() - t0
    ref = checksum_jsonl(want)
    n_want = len(want)
    del want
    out = eng.run_host(text)
    host_ok = out.checksum() == ref and len(out) == n_q and out.n_rows == n_rows
    out.close()
    t = torch.zeros((len(text) + 255) // 128 * 128, dtype=torch.uint8, device="cuda")
    t[:len(text)] = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    out = eng.run_device(t.data_ptr(), len(text), stream)
    dev_ok = out.checksum() == ref
    out.close()
    out = eng.run_device_resident(t.data_ptr(), len(text), stream)
    res_ok = out.download().checksum() == ref
    out.close()
    line = {"shape": a.shape, "queries": n_q, "rows": n_rows, "text_gb": round(len(text) / 1e9, 3), "taxa": sh["taxa"], "strategy": sh["strategy"],
            "oracle_s": round(t_oracle, 1), "oracle_threads": os.cpu_count(), "jsonl_bytes": n_want, "checksum": ref,
            "every_query_equal": {"host_text": host_ok, "device_text": dev_ok, "device_resident_then_download": res_ok}}
    print(json.dumps(line))
    sys.exit(0 if host_ok and dev_ok and res_ok else 3)


if __name__ == "__main__":
    main()
