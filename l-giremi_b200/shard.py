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


def plan(units, world_size):
    """bin_of[unit], load[bin] for `world_size` GPUs."""
    return partition_lpt(unit_costs(units), world_size)


def local_shard(pb: PlaneBatch, rank, world_size):
    """(sub-batch of the units assigned to `rank`, their global indices ascending)."""
    bin_of, _ = plan(pb.units, world_size)
    index = np.nonzero(bin_of == rank)[0]
    return pb.subset(index), index


class MergedResult:
    """Rank 0's view after the gather: same fields as api.StepResult, global order."""

    def __init__(self, records, site_mean, site_cnt, unit_rec_off, n_candidates):
        self.records, self.site_mean, self.site_cnt = records, site_mean, site_cnt
        self.unit_rec_off, self.n_candidates = unit_rec_off, n_candidates
        self.n_records = len(records)

    def unit_records(self, unit):
        return self.records[int(self.unit_rec_off[unit]):int(self.unit_rec_off[unit + 1])]


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
    units = pb_global.units
    n_units = len(units)
    n_sites = pb_global.n_sites
    site_mean = np.full(n_sites, np.nan)
    site_cnt = np.zeros(n_sites, dtype=np.uint32)
    per_unit_count = np.zeros(n_units, dtype=np.int64)
    pieces = []
    for r in range(world):
        recs = gathered[0][r].view(PAIR_REC)
        mean = gathered[1][r].view(np.float64)
        cnt = gathered[2][r].view(np.uint32)
        off = gathered[3][r].view(np.uint64).astype(np.int64)
        index = gathered[4][r].view(np.int64)
        assert len(off) == len(index) + 1
        per_unit_count[index] = np.diff(off)
        local_site = 0
        for k, g in enumerate(index.tolist()):
            s0, ns = int(units['site_off'][g]), int(units['n_sites'][g])
            site_mean[s0:s0 + ns] = mean[local_site:local_site + ns]
            site_cnt[s0:s0 + ns] = cnt[local_site:local_site + ns]
            local_site += ns
        recs = recs.copy()
        recs['unit'] = index[recs['unit']] if len(recs) else recs['unit']
        pieces.append(recs)
    records = np.concatenate(pieces) if pieces else np.zeros(0, PAIR_REC)
    # ranks hold ascending global indices and emit in (unit, i, j) order, so a stable sort by unit suffices
    records = records[np.argsort(records['unit'], kind='stable')]
    unit_rec_off = np.zeros(n_units + 1, dtype=np.uint64)
    unit_rec_off[1:] = np.cumsum(per_unit_count)
    s = units['n_sites'].astype(np.int64)
    return MergedResult(records, site_mean, site_cnt, unit_rec_off, int((s * (s - 1) // 2).sum()))
