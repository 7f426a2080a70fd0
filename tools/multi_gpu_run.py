"""The north-star run through the product's own entry point: ONE hit table in pinned host memory -> N GPUs of one box ->
ONE result, in one process (blu_ctx_create_multi: query-range shards, one host worker thread per GPU, no collective).

    python tools/multi_gpu_run.py --devices 0,1,2,3,4,5,6,7 --queries-per-gpu 1000000 --hits 50

Prints one JSON line: end-to-end queries/s of blu_consensus_run_host on the multi-device context (text H2D + kernels + result
D2H for all shards inside the timed region), the same table through one GPU for comparison, and the parity evidence: the
whole-table checksum of the sharded run == the checksum of the single-GPU run == (for a prefix) the oracle."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0")
    ap.add_argument("--queries-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--hits", type=int, default=50)
    ap.add_argument("--taxa", type=int, default=30_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--text-refs", action="store_true", help="BLU_OPT_TEXT_REFS: strings stay references into the caller's text")
    ap.add_argument("--skip-single", action="store_true")
    args = ap.parse_args()
    devices = [int(d) for d in args.devices.split(",")]
    from blutils_b200 import _ffi, ConsensusEngine, ConsensusStrategy, Taxon
    from blutils_b200.synth import SynthWorkload
    from oracle_ffi import Oracle, checksum_jsonl

    nq = args.queries_per_gpu * len(devices)
    w = SynthWorkload(args.taxa, seed=20261020)
    ids, off, blob = w.lineages()
    cap = nq * args.hits * 84 + (1 << 20)
    pinned = _ffi.lib().blu_host_alloc(cap)
    if not pinned:
        raise MemoryError("pinned allocation failed")
    t0 = time.perf_counter()
    nbytes, nrows = w.hits_into(pinned, cap, 0, nq, args.hits)
    gen_s = time.perf_counter() - t0

    def run(devs, steps):
        eng = ConsensusEngine(Taxon.Bacteria, ConsensusStrategy.Relaxed, devices=devs if len(devs) > 1 else None, device=devs[0],
                              text_refs=args.text_refs)
        eng.load_taxonomy_raw(ids.ctypes.data, off.ctypes.data, blob.ctypes.data, len(ids))
        eng.run_host(pinned, nbytes).close()  # warm-up: allocations, output densities
        best = None
        out = None
        for _ in range(steps):
            if out is not None:
                out.close()
            t0 = time.perf_counter()
            out = eng.run_host(pinned, nbytes)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        tm = eng.timings()
        res = {"devices": devs, "ms": best * 1e3, "queries_per_s": len(out) / best, "text_gb_per_s": nbytes / best / 1e9, "queries": len(out),
               "rows": out.n_rows, "d2h_bytes": int(tm["d2h_bytes"]), "h2d_bytes": int(tm["h2d_bytes"]), "checksum": out.checksum()}
        first = out.jsonl(head=20000)
        out.close()
        eng.close()
        return res, first

    multi, first_multi = run(devices, args.steps)
    line = {"workload": f"one table of {nq} queries x {args.hits} hits ({nbytes / 1e9:.2f} GB of text) in pinned host memory, {args.taxa}-taxon map",
            "generate_s": round(gen_s, 1), "host_cpus": os.cpu_count(), "multi": multi}
    lin = [bytes(blob[int(off[i]):int(off[i + 1])]).decode() for i in range(len(ids))]
    want = Oracle(ids.tolist(), lin, "bacteria", "relaxed", None, threads=os.cpu_count()).run_raw(w.hits(0, 20000, args.hits))[0]
    line["first_20000_queries_equal_oracle"] = first_multi == want
    if not args.skip_single and len(devices) > 1:
        single, first_single = run(devices[:1], max(1, args.steps - 1))
        line["single"] = single
        line["sharded_checksum_equals_single_gpu"] = single["checksum"] == multi["checksum"]
        line["speedup_vs_one_gpu"] = single["ms"] / multi["ms"]
    line["ok"] = bool(line["first_20000_queries_equal_oracle"] and line.get("sharded_checksum_equals_single_gpu", True))
    print(json.dumps(line))
    sys.exit(0 if line["ok"] else 3)


if __name__ == "__main__":
    main()
