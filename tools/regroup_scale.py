#!/usr/bin/env python
"""Wall time of a NON-contiguous hit table (every query's rows in two fragments, half a file apart) through
blu_consensus_run_host with the regrouping done on the GPU (blu_regroup.cu) and on the host (BLU_REGROUP_HOST=1), next to
the contiguous table of the same rows.  Measurement tool, not part of pytest / bench.py.

  python tools/regroup_scale.py --queries 500000 --hits 25
"""
import argparse, ctypes as C, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=500_000)
    ap.add_argument("--hits", type=int, default=25)
    a = ap.parse_args()
    from blutils_b200 import ConsensusEngine, ConsensusStrategy, Taxon, _ffi
    from blutils_b200.synth import SynthWorkload

    w = SynthWorkload(30000, seed=20261018 + 7)
    ids, off, blob = w.lineages(False)
    eng = ConsensusEngine(Taxon.Bacteria, ConsensusStrategy.Relaxed, False, None)
    eng.load_taxonomy_raw(ids.ctypes.data, off.ctypes.data, blob.ctypes.data, len(ids))
    cap = 2 * a.queries * a.hits * 84 + (1 << 20)
    pinned = _ffi.lib().blu_host_alloc(cap)
    assert pinned
    n1, r1 = w.hits_into(pinned, cap, 0, a.queries, a.hits)
    C.memmove(pinned + n1, pinned, n1)  # the same queries again, half a file later
    nbytes = 2 * n1

    def run(label, env):
        if env:
            os.environ["BLU_REGROUP_HOST"] = "1"
        else:
            os.environ.pop("BLU_REGROUP_HOST", None)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); out = eng.run_host(pinned, nbytes); ts.append(time.perf_counter() - t0)
            res = (len(out), out.n_rows, out.checksum(), int(eng.timings()["n_regrouped"]))
            out.close()
        print(json.dumps({"path": label, "ms": round(min(ts) * 1e3, 1), "all_ms": [round(t * 1e3, 1) for t in ts], "gb_per_s": round(nbytes / min(ts) / 1e9, 2), "queries": res[0], "rows": res[1],
                          "n_regrouped": res[3], "checksum": res[2]}), flush=True)
        return res

    print(json.dumps({"table": f"{a.queries} queries x 2 fragments x {a.hits} rows", "bytes": nbytes}), flush=True)
    g = run("scattered, regrouped on the GPU", False)
    h = run("scattered, regrouped on the host", True)
    os.environ.pop("BLU_REGROUP_HOST", None)
    t0 = time.perf_counter(); out = eng.run_host(pinned, n1); dt = time.perf_counter() - t0
    print(json.dumps({"path": "contiguous half (one fragment per query), for scale", "ms": round(dt * 1e3, 1), "gb_per_s": round(n1 / dt / 1e9, 2), "queries": len(out)}))
    out.close()
    ok = g[:3] == h[:3] and g[3] == 2 and h[3] == 1
    print(json.dumps({"gpu_equals_host": ok}))
    sys.exit(0 if ok else 3)


if __name__ == "__main__":
    main()
