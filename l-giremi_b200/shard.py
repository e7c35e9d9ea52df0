"""Multi-GPU sharding of the MI step: one process per GPU, units partitioned by
pre-computed pair-count cost (longest-processing-time bin packing), NO
collective on the data path -- units are independent.  The only
cross-unit step is the global ECDF / mip pass (giremi.py:415-429), which needs
every site's mean MI in one place: per-rank results are gathered to rank 0 and
put back into the reference's row order (unit -> pair).

torch.distributed is plumbing here (gather of result buffers); it works with
the "nccl" backend (device tensors) and with "gloo" (host tensors, used by the
CPU tests)."""
from __future__ import annotations

import numpy as np

from ._lib import PAIR_REC
from .api import partition_lpt, unit_costs
from .encode import PlaneBatch
from .multigpu import MergedResult, merge_shards  # noqa: F401  (MergedResult: rank 0's view after the gather)


def plan(units, world_size):
    """bin_of[unit], load[bin] for `world_size` GPUs."""
    return partition_lpt(unit_costs(units), world_size)


def local_shard(pb: PlaneBatch, rank, world_size):
    """(sub-batch of the units assigned to `rank`, their global indices ascending)."""
    bin_of, _ = plan(pb.units, world_size)
    index = np.nonzero(bin_of == rank)[0]
    return pb.subset(index), index


def _gather_bytes(buf: np.ndarray, dst, group, device):
    """Gathers one variable-length byte buffer per rank onto `dst` (list of
    numpy arrays there, None elsewhere)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([buf.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    mine = torch.zeros(cap, dtype=torch.uint8, device=device)
    if buf.size:
        mine[:buf.size] = torch.from_numpy(buf.copy()).to(device)
    parts = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(mine, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return [p[:s].cpu().numpy() for p, s in zip(parts, sizes)]


def gather_to_rank0(pb_global: PlaneBatch, local_result, local_index, group=None, device="cpu"):
    """Collects every rank's StepResult on rank 0 and restores global order.

    pb_global is only used for its unit table (site offsets / counts); every
    rank holds it (descriptors are tiny), only rank 0 gets a MergedResult."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    rec = np.ascontiguousarray(local_result.records)
    payloads = [
        rec.view(np.uint8).reshape(-1),
        np.ascontiguousarray(local_result.site_mean, dtype=np.float64).view(np.uint8).reshape(-1),
        np.ascontiguousarray(local_result.site_cnt, dtype=np.uint32).view(np.uint8).reshape(-1),
        np.ascontiguousarray(local_result.unit_rec_off, dtype=np.uint64).view(np.uint8).reshape(-1),
        np.ascontiguousarray(local_index, dtype=np.int64).view(np.uint8).reshape(-1),
    ]
    gathered = [_gather_bytes(p, 0, group, device) for p in payloads]
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        off = gathered[3][r].view(np.uint64)
        index = gathered[4][r].view(np.int64)
        assert len(off) == len(index) + 1
        parts.append((index, gathered[0][r].view(PAIR_REC), gathered[1][r].view(np.float64),
                      gathered[2][r].view(np.uint32), off, None))
    # ranks hold ascending global indices and emit in (unit, i, j) order: the merge is index bookkeeping
    return merge_shards(pb_global.units, pb_global.n_sites, parts)
