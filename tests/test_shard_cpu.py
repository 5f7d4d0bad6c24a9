"""Multi-process (world_size 2, gloo) tests of the sharding logic used by bench.py / an embedding program at N > 1:
contiguous cell blocks, one SUM all-reduce of the spectra, concatenation of per-rank event lists.  The per-rank
compute stand-in here is the CPU oracle (test infrastructure); on the GPU box the same helpers run over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import harness
import oracle_api
from is3d2_b200 import shard, workdir


def test_cell_range_tiles_the_surface():
    for n in (0, 1, 7, 8, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard.cell_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.cell_range(10, 2, 2)


def test_merge_event_lists_groups_by_event():
    dt = np.dtype([("event", "<i4"), ("tag", "<i4")])
    a = np.array([(0, 10), (0, 11), (2, 12)], dtype=dt)
    b = np.array([(1, 20), (2, 21), (2, 22)], dtype=dt)
    out, counts = shard.merge_event_lists([(a, [2, 0, 1]), (b, [0, 1, 2])], 3)
    assert counts.tolist() == [2, 1, 3]
    assert out["event"].tolist() == [0, 0, 1, 2, 2, 2]
    assert out["tag"].tolist() == [10, 11, 20, 12, 21, 22]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, tmp: str, name: str, queue):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = cases.SPECTRA_CASES[name]
        surf, _ = harness.load_golden(name)
        mine, offset = shard.shard_surface(surf, rank, world)
        root = workdir.make_workdir(os.path.join(tmp, f"rank{rank}"), case["params"], chosen=case["chosen"])
        prob = oracle_api.OracleProblem(root, case["params"], {k: np.ascontiguousarray(v) for k, v in mine.items()},
                                        after_surface=shard.set_global_thermo_averages)
        rc, part, _ = prob.spectra()
        assert rc == 0
        t = torch.from_numpy(np.ascontiguousarray(part))
        shard.allreduce_sum_(t)
        rc, ntot = prob.total_yield() if case["params"]["df_mode"] != 5 else (0, 0.0)
        y = torch.tensor([ntot], dtype=torch.float64)
        shard.allreduce_sum_(y)
        if rank == 0:
            queue.put((t.numpy().copy(), float(y.item()), offset))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["s3d_m2_baryon", "s3d_m3"])
def test_two_rank_allreduce_matches_whole_surface(libs, tmp_path, name):
    """world_size 2 over gloo: per-rank spectra of the two cell blocks, SUM all-reduced, equal the reference's
    golden spectra of the whole surface (the blocks change only the summation order)."""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), name, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got, ntot, offset0 = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert offset0 == 0
    _, ref = harness.load_golden(name)
    harness.assert_spectra_close(got, ref, rtol=1e-11, what=f"{name} 2-rank")
    # the total yield is additive over the blocks as well
    case = cases.SPECTRA_CASES[name]
    surf, _ = harness.load_golden(name)
    root = workdir.make_workdir(str(tmp_path / "whole"), case["params"], chosen=case["chosen"])
    rc, whole = oracle_api.OracleProblem(root, case["params"], surf).total_yield()
    assert rc == 0 and abs(ntot / whole - 1.0) < 1e-12
