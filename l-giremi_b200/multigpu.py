"""One host process driving every GPU of the box ("CUDA context only in the
parent"; north_star part 4).

The reference's CLI forks `-t` workers that each run the whole per-region analysis
(giremi.py:375-380).  With the MI step on GPUs the workers only extract and encode; the PARENT
owns the devices: it partitions the (footprint, strand) units of all chunks by pre-computed
pair-count cost (longest-processing-time bin packing, lgmi_partition_lpt), submits one shard per
GPU from a thread pool (the C-ABI calls release the GIL; a handle is used by one thread at a time),
and merges the results back into the reference's row order.  Units are independent:
no collective, no peer traffic.  `shard.py` is the same thing for one PROCESS per GPU
(torch.distributed launches: bench.py --gpus N)."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import api
from ._lib import PAIR_REC
from .encode import PlaneBatch


class MergedResult:
    """Results of several shards put back in global unit order: the fields of api.StepResult
    that callers read (records, site_mean, site_cnt, unit_rec_off, counts)."""

    def __init__(self, records, site_mean, site_cnt, unit_rec_off, n_candidates, counts=None, shard_ms=None,
                 shard_units=None):
        self.records, self.site_mean, self.site_cnt = records, site_mean, site_cnt
        self.unit_rec_off, self.n_candidates, self.counts = unit_rec_off, n_candidates, counts
        self.n_records = len(records)
        self.shard_ms, self.shard_units = shard_ms, shard_units

    def unit_records(self, unit):
        return self.records[int(self.unit_rec_off[unit]):int(self.unit_rec_off[unit + 1])]

    def unit_counts(self, unit):
        if self.counts is None:
            return np.zeros((0, 9), dtype=np.uint32)
        return self.counts[int(self.unit_rec_off[unit]):int(self.unit_rec_off[unit + 1])]


def merge_shards(units, n_sites, parts) -> MergedResult:
    """`parts` = [(global unit indices ascending, records, site_mean, site_cnt, unit_rec_off, counts | None)]
    per shard, each in its own local order -> one result in global order.  Pure index bookkeeping."""
    n_units = len(units)
    site_mean = np.full(n_sites, np.nan)
    site_cnt = np.zeros(n_sites, dtype=np.uint32)
    per_unit = np.zeros(n_units, dtype=np.int64)
    site_off = units['site_off'].astype(np.int64)
    n_s = units['n_sites'].astype(np.int64)
    for index, _rec, mean, cnt, off, _counts in parts:
        index = np.asarray(index, dtype=np.int64)
        per_unit[index] = np.diff(np.asarray(off).astype(np.int64))
        # the shard's sites are its units' sites back to back, in shard order
        ns = n_s[index]
        if ns.sum():
            dst = np.repeat(site_off[index] - np.concatenate(([0], np.cumsum(ns)[:-1])), ns) + np.arange(int(ns.sum()))
            site_mean[dst] = np.asarray(mean)[:len(dst)]
            site_cnt[dst] = np.asarray(cnt)[:len(dst)]
    unit_rec_off = np.zeros(n_units + 1, dtype=np.uint64)
    unit_rec_off[1:] = np.cumsum(per_unit)
    total = int(unit_rec_off[-1])
    records = np.empty(total, dtype=PAIR_REC)
    want_counts = any(p[5] is not None for p in parts)
    counts = np.empty((total, 9), dtype=np.uint32) if want_counts else None
    for index, rec, _mean, _cnt, off, cnts in parts:
        index = np.asarray(index, dtype=np.int64)
        if not len(index):
            continue
        off = np.asarray(off).astype(np.int64)
        n = np.diff(off)
        # row r of the shard (unit k = its position in the shard) goes to unit_rec_off[index[k]] + (r - off[k])
        shift = np.repeat(unit_rec_off[index].astype(np.int64) - off[:-1], n)
        dst = shift + np.arange(int(n.sum()))
        rec = np.asarray(rec)
        records[dst] = rec
        records['unit'][dst] = np.repeat(index, n).astype(np.uint32)
        if counts is not None and cnts is not None and len(dst):
            counts[dst] = cnts
    return MergedResult(records, site_mean, site_cnt, unit_rec_off, int((n_s * (n_s - 1) // 2).sum()), counts)


class DevicePool:
    """Contexts on `devices` (default: every visible GPU), owned by the calling process."""

    def __init__(self, devices=None):
        if devices is None:
            import os
            env = os.environ.get("LGMI_DEVICES", "")                  # e.g. "0,1,2,3"; default: every visible GPU
            devices = [int(d) for d in env.split(",") if d != ""] or list(range(max(1, api.device_count())))
        self.devices = [int(d) for d in devices]
        self.contexts = [api.Context(d) for d in self.devices]       # raises without a GPU: no CPU path
        self._threads = ThreadPoolExecutor(max_workers=len(self.devices)) if len(self.devices) > 1 else None

    def __len__(self):
        return len(self.devices)

    def close(self):
        if self._threads is not None:
            self._threads.shutdown()
            self._threads = None
        for c in self.contexts:
            c.close()
        self.contexts = []

    def plan(self, pb: PlaneBatch):
        """[ascending global unit indices of each device's shard] by LPT on S(S-1)/2 * ceil(R/64)."""
        bin_of, _load = api.partition_lpt(api.unit_costs(pb.units), len(self.devices))
        return [np.flatnonzero(bin_of == k) for k in range(len(self.devices))]

    def run(self, pb: PlaneBatch, min_common_reads=5, mode=api.MODE_HET_ONLY):
        """The MI step of `pb` over all devices; same result as api.mi_step_batched on one."""
        if len(self.devices) == 1 or pb.n_units < 2 * len(self.devices):
            return api.mi_step_batched(pb, min_common_reads, mode, ctx=self.contexts[0])
        shards = self.plan(pb)

        def work(k):
            index = shards[k]
            if not len(index):
                return None
            import time
            t0 = time.perf_counter()
            res = api.mi_step_batched(pb.subset(index), min_common_reads, mode, ctx=self.contexts[k])
            return (index, res.records, res.site_mean, res.site_cnt, res.unit_rec_off, res.counts,
                    1e3 * (time.perf_counter() - t0))

        done = [r for r in self._threads.map(work, range(len(self.devices))) if r is not None]
        merged = merge_shards(pb.units, pb.n_sites, [r[:6] for r in done])
        merged.shard_ms = [r[6] for r in done]
        merged.shard_units = [len(r[0]) for r in done]
        return merged


_pool = None


def get_pool() -> DevicePool:
    """Process-wide pool over every visible GPU, created on first use (in the parent, after the
    extraction workers have been forked and have returned)."""
    global _pool
    if _pool is None or not _pool.contexts:
        _pool = DevicePool()
    return _pool
