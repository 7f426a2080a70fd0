#!/usr/bin/env python
"""bench.py -- consensus queries/s of the blutils consensus-identity hot path on B200.

Workload at N=1 (BASELINE.json configs[1], "C2"): synthetic 16S amplicon run, 1M queries x 50 hits, lineage map of
30k taxa, `--taxon custom` with the cutoffs of assets/custom-taxon-cutoffs-bacteria-16S.yaml, strategy relaxed.
N>1: every rank processes its own query range of the same shape (queries are independent -> weak scaling,
no data-path collective).

One "step" = one pass of the hot path over the rank's hit table.
  value : device-resident throughput (text already in HBM; kernels + result download), CUDA events, max over ranks
  e2e   : the same through the reference-facing C-ABI call with the text in pinned HOST memory
          (chunked H2D + kernels + result D2H inside the timed region)
  roofline : the fused tile kernel; achieved = algorithmic bytes (text + result records + lineage tables) / its
          CUDA-event duration measured in this run, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the C++ oracle port of the reference timed on the host cores on a bounded sample of the workload
`--impl reference` times that CPU port alone (the reference is Rust + polars and cannot be built in this image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q_PER_GPU = 1_000_000
HITS = 50
N_TAXA = 30_000
SEED = 20261018 + 2  # base seed + config id
CUSTOM = {"domain": 50, "kingdom": 60, "phylum": 75, "class": 80, "order": 85, "family": 92, "genus": 97, "species": 99}
CPU_SAMPLE_Q = 100_000


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self._t = None

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self._t = threading.Thread(target=self._read, daemon=True)
        self._t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for i, n in enumerate(names):
                if len(r) > 4 + i and r[4 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(rank: int, q_per_gpu: int, hits: int, n_taxa: int):
    """Synthetic lineage map + this rank's hit table, generated straight into pinned host memory."""
    from blutils_b200 import _ffi
    from blutils_b200.synth import SynthWorkload

    w = SynthWorkload(n_taxa, seed=SEED)
    ids, off, blob = w.lineages(numeric=False)
    cap = q_per_gpu * hits * 84 + (1 << 20)
    pinned = _ffi.lib().blu_host_alloc(cap)
    if not pinned:
        raise MemoryError("pinned allocation failed")
    nbytes, nrows = w.hits_into(pinned, cap, rank * q_per_gpu, q_per_gpu, hits)
    return w, (ids, off, blob), pinned, nbytes, nrows


def cpu_reference(w, lineages, q_sample: int, hits: int, steps: int, warmup: int, q_begin: int = 0):
    """Times the CPU port of the reference (oracle/blu_oracle.cpp, all host threads) on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle_ffi import Oracle

    ids, off, blob = lineages
    lin = [bytes(blob[int(off[i]):int(off[i + 1])]).decode() for i in range(len(ids))]
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))  # the CPU arm may use every core of the box
    except OSError:
        pass
    cores = len(os.sched_getaffinity(0))
    orc = Oracle(ids.tolist(), lin, "custom", "relaxed", CUSTOM, threads=cores)
    text = w.hits(q_begin, q_sample, hits)
    times = []
    nq = nr = 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, nq, nr = orc.run_raw(text)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return {"value": nq / t, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{q_sample} queries x {hits} hits ({len(text) / 1e6:.0f} MB text), {len(times)} pass(es), {cores} threads; "
                      f"C++ restatement of blutils 8.3.1 (Rust+polars reference not buildable here)",
            "rows_per_s": nr / t, "ms_per_pass": t * 1e3}, t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=Q_PER_GPU, help="queries per GPU (default = BASELINE config C2)")
    ap.add_argument("--hits", type=int, default=HITS)
    ap.add_argument("--taxa", type=int, default=N_TAXA)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    config = {"workload": f"C2 synthetic 16S amplicon run: {args.queries} queries x {args.hits} hits per GPU, {args.taxa}-taxon lineage map, "
                          "--taxon custom (custom-taxon-cutoffs-bacteria-16S.yaml values), strategy relaxed",
              "queries_per_gpu": args.queries, "hits_per_query": args.hits, "taxa": args.taxa, "sharding": f"query-range x{world}, no collective",
              "l2_policy": "input (GBs) larger than the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        from blutils_b200.synth import SynthWorkload

        w = SynthWorkload(args.taxa, seed=SEED)
        sample = min(args.queries, CPU_SAMPLE_Q)
        base, t = cpu_reference(w, w.lineages(False), sample, args.hits, args.steps, min(args.warmup, 1))
        line = {"impl": "reference", "metric": "consensus_queries_per_s", "value": base["value"], "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config, "cpu_baseline": base,
                "hit_rows_per_s": base["rows_per_s"],
                "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    from blutils_b200 import ConsensusEngine, ConsensusStrategy, CustomTaxon, Taxon

    torch.cuda.set_device(local_rank)
    # keep this rank's threads and its pinned staging memory on the NUMA node of its GPU (one PCIe link per GPU)
    try:
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        config["cpu_affinity"] = f"{len(os.sched_getaffinity(0))} cpus (NVML ideal affinity of GPU {local_rank})"
    except Exception as ex:  # noqa: BLE001
        config["cpu_affinity"] = f"not set ({type(ex).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w, lineages, pinned, nbytes, nrows = make_workload(rank, args.queries, args.hits, args.taxa)
    ids, off, blob = lineages

    # custom cutoffs go through the YAML path (CustomTaxon::from_file semantics)
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        for k, v in CUSTOM.items():
            f.write(f"{k}: {v}\n")
        ypath = f.name
    custom = CustomTaxon.from_file(ypath)
    os.unlink(ypath)
    eng = ConsensusEngine(Taxon.Custom, ConsensusStrategy.Relaxed, False, custom, device=local_rank)
    eng.load_taxonomy_raw(ids.ctypes.data, off.ctypes.data, blob.ctypes.data, len(ids))

    # device-resident copy of the text
    dbuf = torch.empty((nbytes + 255) // 128 * 128, dtype=torch.uint8, device="cuda")
    host_view = (C.c_uint8 * nbytes).from_address(pinned)
    dbuf[:nbytes].copy_(torch.frombuffer(host_view, dtype=torch.uint8))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    # ---------------- device-resident arm -------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # spans warm-up + timed region; the median is taken over samples under load
    for _ in range(args.warmup):
        eng.run_device(dbuf.data_ptr(), nbytes, stream).close()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tile_ms, long_ms, gather_ms, launches = [], [], [], 0
    tile_launches = 1
    nq = 0
    res_bytes = tax_bytes = 0
    e0.record()
    for _ in range(args.steps):
        out = eng.run_device(dbuf.data_ptr(), nbytes, stream)
        t = eng.timings()
        tile_ms.append(t["ms_tile_kernel"])
        long_ms.append(t["ms_longrun_kernel"])
        gather_ms.append(t["ms_gather_kernel"])
        launches += int(t["n_kernel_launches"])
        tile_launches = max(1, int(t["n_tile_launches"]))
        nq = len(out)
        res_bytes, tax_bytes = int(t["result_bytes"]), int(t["taxonomy_bytes"])
        out.close()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    total_q = sum_over_ranks(float(nq))
    total_rows = sum_over_ranks(float(nrows))

    # ---------------- end-to-end arm: pinned host text through the C ABI ------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        eng.run_host(pinned, nbytes).close()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        out = eng.run_host(pinned, nbytes)
        d2h = int(eng.timings()["d2h_bytes"])
        out.close()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    barrier()
    h2d_gbps = eng.measure_h2d(1 << 30) if rank == 0 else 0.0

    # ---------------- roofline of the dominant kernel ---------------------------------------------------------------
    peaks = {}
    for pth in (os.path.join(ROOT, "MEASURED_PEAKS.json"),):
        if os.path.exists(pth):
            peaks = json.load(open(pth))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "B200_PROFILING.md fallback 6650 GB/s"
    algo_bytes = nbytes + res_bytes + tax_bytes  # B_text + B_out + B_tax per GPU
    # DRAM traffic of one launch from the committed `ncu --set full` capture of this same workload (else null)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_tile_kernel_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if int(tj.get("text_bytes", -1)) == int(nbytes) and int(tj.get("launches_per_step", 1)) == tile_launches:
            traffic = int(tj["dram_read_bytes"] + tj["dram_write_bytes"])  # of ONE launch, like `achieved`
    tile_avg = sum(tile_ms) / len(tile_ms)
    achieved = algo_bytes / (tile_avg * 1e-3) / 1e9 if tile_avg > 0 else 0.0

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference(w, lineages, min(args.queries, CPU_SAMPLE_Q), args.hits, 1, 1)

    if rank == 0:
        line = {
            "metric": "consensus_queries_per_s", "value": total_q / (dev_ms * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config,
            "hit_rows_per_s": total_rows / (dev_ms * 1e-3),
            "text_gb_per_s": world * nbytes / (dev_ms * 1e-3) / 1e9,
            "e2e": {"value": total_q / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "text_gb_per_s": world * nbytes / e2e_s / 1e9, "pinned_h2d_peak_gb_per_s": h2d_gbps,
                    "hit_rows_per_s": total_rows / e2e_s},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "tile_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(algo_bytes // tile_launches),
                         "ms_per_launch": tile_avg / tile_launches, "launches_per_step": tile_launches,
                         "ms_per_step_all_launches": tile_avg, "kernel_share_of_step": tile_avg / dev_ms if dev_ms else None,
                         "ms_longrun_kernel": sum(long_ms) / len(long_ms), "ms_gather_dup_kernels": sum(gather_ms) / len(gather_ms)},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
