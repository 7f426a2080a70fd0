"""N>1 host logic on CPU (world_size 2, gloo): every rank computes the same query-aligned cuts with
blu_shard_cuts, processes only its own byte range and the ranks' results are gathered without any data-path
collective (an object gather of the finished JSONL stands in for the D2H gather of result buffers).
The per-shard consensus itself runs through the host simulation of the device logic (no GPU in this container);
the GPU equivalent is tests/test_gpu_parity.py::test_sharded_equals_whole."""
import os
import random
import socket
import sys

import pytest

from helpers import ROOT, random_blast, random_taxonomy


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(seed):
    rng = random.Random(seed)
    units = random_taxonomy(rng, n_leaves=40)
    return units, random_blast(rng, units, n_queries=120, contiguous=True)


def _worker(rank, world, port, out_dir, seed):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import sim_ffi
    from blutils_b200 import shard_cuts

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units, text = _case(seed)
    ids = [u["taxid"] for u in units]
    lin = [u["textLineage"] for u in units]
    cuts = shard_cuts(text, world)
    a, b = cuts[rank], cuts[rank + 1]
    rc, js, err = sim_ffi.run(ids, lin, "fungi", "cautious", text[a:b]) if b > a else (0, b"", "")
    assert rc == 0, err
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, a, b, js))
    if rank == 0:
        from oracle_ffi import Oracle

        want = Oracle(ids, lin, "fungi", "cautious").run_raw(text)[0]
        lines = []
        covered = 0
        for r, sa, sb, part in sorted(gathered):
            assert sa == covered
            covered = sb
            lines += part.decode().splitlines()
        assert covered == len(text)
        got = ("\n".join(sorted(lines, key=lambda l: l.encode())) + "\n").encode()
        open(os.path.join(out_dir, "ok"), "w").write("1" if got == want else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(tmp_path):
    import torch.multiprocessing as mp

    # the cuts never split a query
    from blutils_b200 import shard_cuts

    rng = random.Random(5)
    units = random_taxonomy(rng, n_leaves=10)
    text = random_blast(rng, units, n_queries=50, contiguous=True)
    for n in (1, 2, 5, 16, 200):
        cuts = shard_cuts(text, n)
        assert cuts[0] == 0 and cuts[-1] == len(text) and cuts == sorted(cuts)
        for c in cuts[1:-1]:
            if 0 < c < len(text):
                assert text[c - 1:c] == b"\n"
                prev = text[:c - 1].rsplit(b"\n", 1)[-1].split(b"\t")[0]
                assert text[c:].split(b"\t", 1)[0] != prev
    # a generated case the reference accepts (no abort input), found deterministically
    from oracle_ffi import Oracle, OracleDataError

    seed = None
    for cand in range(99, 140):
        u, t = _case(cand)
        try:
            Oracle([x["taxid"] for x in u], [x["textLineage"] for x in u], "fungi", "cautious").run_raw(t)
            seed = cand
            break
        except OracleDataError:
            continue
    assert seed is not None
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), seed), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
