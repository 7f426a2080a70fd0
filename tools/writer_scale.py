"""Result writer at scale (write_blutils_output.rs:87-111): time of blu_result_write (JSON pretty, JSONL) and of
the build-tabular emitter next to the consensus step that produced the result.  Usage (GPU box):
    python tools/writer_scale.py --queries 1000000 --hits 50
    python tools/writer_scale.py --queries 10000000 --hits 5"""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--hits", type=int, default=50)
    ap.add_argument("--taxa", type=int, default=30_000)
    ap.add_argument("--out-dir", default="/dev/shm")
    args = ap.parse_args()
    from blutils_b200 import _ffi, ConsensusEngine, ConsensusStrategy, OutputFormat, Taxon
    from blutils_b200.synth import SynthWorkload

    w = SynthWorkload(args.taxa, seed=20261020)
    ids, off, blob = w.lineages()
    cap = args.queries * args.hits * 84 + (1 << 20)
    pinned = _ffi.lib().blu_host_alloc(cap)
    nbytes, nrows = w.hits_into(pinned, cap, 0, args.queries, args.hits)
    eng = ConsensusEngine(Taxon.Bacteria, ConsensusStrategy.Relaxed)
    eng.load_taxonomy_raw(ids.ctypes.data, off.ctypes.data, blob.ctypes.data, len(ids))
    eng.run_host(pinned, nbytes).close()
    t0 = time.perf_counter()
    out = eng.run_host(pinned, nbytes)
    t_run = time.perf_counter() - t0
    print(f"{args.queries} queries x {args.hits} hits: {nbytes / 1e9:.2f} GB text, consensus (run_host, PCIe-bound) {t_run * 1e3:.1f} ms, "
          f"{os.cpu_count()} host cpus")
    for name, fmt in (("jsonl", OutputFormat.Jsonl), ("json", OutputFormat.Json), ("yaml", OutputFormat.Yaml)):
        path = os.path.join(args.out_dir, f"blu_writer_scale.{name}")
        t0 = time.perf_counter()
        out.write(path, fmt, run_id="00000000-0000-4000-8000-000000000000")
        dt = time.perf_counter() - t0
        size = os.path.getsize(path)
        print(f"  write {name:5s}: {dt * 1e3:9.1f} ms  {size / 1e9:6.2f} GB  {size / dt / 1e9:5.2f} GB/s  {args.queries / dt / 1e6:6.2f} M queries/s")
        os.unlink(path)
    path = os.path.join(args.out_dir, "blu_writer_scale.tsv")
    t0 = time.perf_counter()
    out.write_tabular(path, run_id="00000000-0000-4000-8000-000000000000")
    dt = time.perf_counter() - t0
    print(f"  write tsv  : {dt * 1e3:9.1f} ms  {os.path.getsize(path) / 1e9:6.2f} GB")
    os.unlink(path)
    out.close()


if __name__ == "__main__":
    main()
