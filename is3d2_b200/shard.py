"""Host-side sharding of the freezeout surface across GPUs: cells are independent additive
contributions, so each rank takes one contiguous block of the structure-of-arrays columns, the tables are replicated,
and the only exchange is one SUM all-reduce of the spectra (or dN/dX histograms, or the scalar total yield).  The
sampler needs no collective: its Philox streams are keyed by the GLOBAL cell index (`global_offset` of
is3d_set_surface), per-rank particle lists are concatenated event by event.

Works with any torch.distributed backend: NCCL on the GPUs (bench.py), gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def cell_range(n_cells: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of rank's contiguous block; blocks differ in size by at most one cell and tile [0, n_cells)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(int(n_cells), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_surface(surface: dict, rank: int, world: int) -> tuple[dict, int]:
    """(rank's block of every column as views, global_offset of its first cell)"""
    n = len(next(iter(surface.values())))
    b, e = cell_range(n, rank, world)
    return {k: v[b:e] for k, v in surface.items()}, b


def allreduce_sum_(tensor):
    """In-place SUM all-reduce of a spectra / histogram / yield tensor over the default process group (no-op when
    torch.distributed is not initialised, i.e. single GPU).  The tensor lives where the backend wants it: CUDA for
    NCCL, CPU for gloo."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def set_global_thermo_averages(session) -> np.ndarray:
    """All-reduce the six additive ds_max-weighted sums of this rank's cell block and write the WHOLE surface's
    averages (T, E, P, muB, nB) into the session's side file; call between set_surface() and prepare().  Fast-mode
    species densities, the PTM fast breakdown test and the PTB tables are evaluated at these averages
    (reference readindata.cpp:330-366, DeltafData.cpp:220-295, :555-690)."""
    import torch
    sums = torch.from_numpy(session.thermo_sums())
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            dev = sums.cuda()
            dist.all_reduce(dev)
            sums = dev.cpu()
        else:
            dist.all_reduce(sums)
    avg = (sums[:5] / sums[5]).numpy()
    session.set_thermo_averages(avg)
    return avg


def merge_event_lists(per_rank: list[tuple[np.ndarray, np.ndarray]], nevents: int) -> tuple[np.ndarray, np.ndarray]:
    """Concatenate per-rank sampler outputs (particles grouped by event, counts per event) into one list grouped by
    event: event e holds rank 0's hadrons of e, then rank 1's, ... -- the order the reference's cell loop would
    produce for contiguous cell blocks (ParticleSampler.cpp:1093-1120 appends per cell)."""
    counts = np.zeros(nevents, dtype=np.int64)
    for _, c in per_rank:
        counts += np.asarray(c, dtype=np.int64)
    dtype = per_rank[0][0].dtype
    out = np.empty(int(counts.sum()), dtype=dtype)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    fill = starts.copy()
    for parts, c in per_rank:
        c = np.asarray(c, dtype=np.int64)
        src = np.concatenate([[0], np.cumsum(c)[:-1]])
        for e in np.nonzero(c)[0]:
            out[fill[e]:fill[e] + c[e]] = parts[src[e]:src[e] + c[e]]
            fill[e] += c[e]
    return out, counts
