#!/usr/bin/env python
"""Wall time of `blu_consensus_run_file` (file -> results in host memory) on the C2 table written to a file, for several
reader-thread counts, next to the in-memory `run_host` of the same bytes.  Measurement tool, not part of pytest / bench.py.

  python tools/file_stream.py --dir /dev/shm --threads 1 4 8 16
"""
import argparse, ctypes as C, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--threads", type=int, nargs="+", default=[1, 4, 8, 16])
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    from blutils_b200 import ConsensusEngine, ConsensusStrategy, Taxon, _ffi
    from blutils_b200.synth import SynthWorkload

    w = SynthWorkload(30000, seed=20261018 + 2)
    ids, off, blob = w.lineages(False)
    eng = ConsensusEngine(Taxon.Bacteria, ConsensusStrategy.Relaxed, False, None)
    eng.load_taxonomy_raw(ids.ctypes.data, off.ctypes.data, blob.ctypes.data, len(ids))
    cap = a.queries * 50 * 80 + (1 << 20)
    pinned = _ffi.lib().blu_host_alloc(cap)
    assert pinned
    nbytes, nrows = w.hits_into(pinned, cap, 0, a.queries, 50)
    path = os.path.join(a.dir, "blu_c2_blast.out")
    with open(path, "wb") as f:
        f.write((C.c_uint8 * nbytes).from_address(pinned))
    try:
        ts = []
        for _ in range(a.steps + 1):
            t0 = time.perf_counter(); out = eng.run_host(pinned, nbytes); ts.append(time.perf_counter() - t0)
            ref = out.checksum(); out.close()
        print(json.dumps({"path": "run_host (pinned memory)", "ms": round(min(ts[1:]) * 1e3, 2), "gb_per_s": round(nbytes / min(ts[1:]) / 1e9, 2)}), flush=True)
        for th in a.threads:
            os.environ["BLU_READ_THREADS"] = str(th)
            ts = []
            for _ in range(a.steps + 1):
                t0 = time.perf_counter(); out = eng.run_file(path); ts.append(time.perf_counter() - t0)
                assert out.checksum() == ref and len(out) == a.queries
                out.close()
            print(json.dumps({"path": "run_file", "read_threads": th, "first_ms": round(ts[0] * 1e3, 2), "ms": round(min(ts[1:]) * 1e3, 2),
                              "gb_per_s": round(nbytes / min(ts[1:]) / 1e9, 2), "queries_per_s": round(a.queries / min(ts[1:]))}), flush=True)
    finally:
        os.remove(path)


if __name__ == "__main__":
    main()
