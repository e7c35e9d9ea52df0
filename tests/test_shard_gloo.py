"""world_size-2 gloo test of the multi-GPU host logic (l-giremi_b200/shard.py):
LPT sharding, per-rank step, gather to rank 0, restoration of the reference's
row order.  No GPU here, so each rank's "device step" is played by the oracle
(test infrastructure); what is under test is everything around it."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeResult:
    pass


def oracle_step(pb, labels, min_common):
    """StepResult-shaped output for a PlaneBatch whose label matrices are known."""
    import c_oracle
    L = importlib.import_module("l-giremi_b200._lib")
    recs, means, cnts, off = [], [], [], [0]
    for u in range(pb.n_units):
        lab = labels[u].astype(np.int16)
        lab[lab == 255] = -1
        S = lab.shape[0]
        if S >= 2:
            i, j, mi, _ = c_oracle.unit_pairs_from_labels(lab.astype(np.int8), None, min_common)
        else:
            i = j = np.zeros(0, np.int32)
            mi = np.zeros(0)
        is_het = ((pb.site_flags[int(pb.units['site_off'][u]):][:S] & 3) == 2).astype(np.uint8)
        keep = (is_het[i] | is_het[j]).astype(bool) if len(i) else np.zeros(0, bool)
        mean, cnt = c_oracle.site_means(S, is_het, i, j, mi) if S else (np.zeros(0), np.zeros(0, np.int32))
        r = np.zeros(int(keep.sum()), dtype=L.PAIR_REC)
        r['unit'], r['i'], r['j'], r['mi'] = u, i[keep], j[keep], mi[keep]
        recs.append(r)
        means.append(mean)
        cnts.append(cnt.astype(np.uint32))
        off.append(off[-1] + len(r))
    out = FakeResult()
    out.records = np.concatenate(recs) if recs else np.zeros(0, L.PAIR_REC)
    out.site_mean = np.concatenate(means) if means else np.zeros(0)
    out.site_cnt = np.concatenate(cnts) if cnts else np.zeros(0, np.uint32)
    out.unit_rec_off = np.array(off, dtype=np.uint64)
    return out


def make_world():
    synth = importlib.import_module("l-giremi_b200.synth")
    pb, raw = synth.make_heavy_tail(4242, 37, keep_raw=True, s_max=60, r_max=400)
    labels = [synth.labels_from_alleles(a) for a, _ in raw]
    return pb, labels


def worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("l-giremi_b200.shard")
    pb, labels = make_world()
    local, index = shard.local_shard(pb, rank, world)
    res = oracle_step(local, [labels[g] for g in index], 6)
    merged = shard.gather_to_rank0(pb, res, index, device="cpu")
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), records=merged.records, site_mean=merged.site_mean,
                 site_cnt=merged.site_cnt, unit_rec_off=merged.unit_rec_off, n_candidates=merged.n_candidates)
    else:
        assert merged is None
    np.save(os.path.join(out_dir, "index%d.npy" % rank), index)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_and_gather_equals_single_process(tmp_path, lg):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    pb, labels = make_world()
    want = oracle_step(pb, labels, 6)
    got = np.load(tmp_path / "merged.npz")
    assert np.array_equal(got["records"], want.records)            # same rows, same order, same bits
    assert np.array_equal(got["site_mean"], want.site_mean, equal_nan=True)
    assert np.array_equal(got["site_cnt"], want.site_cnt)
    assert np.array_equal(got["unit_rec_off"], want.unit_rec_off)
    assert int(got["n_candidates"]) == pb.n_candidates
    i0, i1 = np.load(tmp_path / "index0.npy"), np.load(tmp_path / "index1.npy")
    assert sorted(i0.tolist() + i1.tolist()) == list(range(pb.n_units))
    # cost balance: LPT bound
    cost = lg.unit_costs(pb.units).astype(np.int64)
    assert abs(int(cost[i0].sum()) - int(cost[i1].sum())) <= int(cost.max())


def test_single_rank_shard_is_identity(lg):
    shard = importlib.import_module("l-giremi_b200.shard")
    pb, _ = make_world()
    local, index = shard.local_shard(pb, 0, 1)
    assert index.tolist() == list(range(pb.n_units))
    assert np.array_equal(local.planes, pb.planes) and np.array_equal(local.units, pb.units)


def test_merge_shards_in_process_three_shards_one_empty(lg):
    """multigpu.merge_shards (what the parent-owned DevicePool and the gather both use) on three
    shards, one of them empty, with the 3x3 tables riding along."""
    mg = importlib.import_module("l-giremi_b200.multigpu")
    pb, labels = make_world()
    want = oracle_step(pb, labels, 6)
    bin_of, _ = lg.partition_lpt(lg.unit_costs(pb.units), 2)
    parts = []
    for k in (0, 1, 2):
        index = np.flatnonzero(bin_of == k)                 # shard 2 is empty
        res = oracle_step(pb.subset(index), [labels[g] for g in index], 6)
        counts = np.repeat(np.arange(len(res.records), dtype=np.uint32)[:, None] * 10 + k, 9, axis=1)
        parts.append((index, res.records, res.site_mean, res.site_cnt, res.unit_rec_off, counts))
    got = mg.merge_shards(pb.units, pb.n_sites, parts)
    assert np.array_equal(got.records, want.records)
    assert np.array_equal(got.site_mean, want.site_mean, equal_nan=True) and np.array_equal(got.site_cnt, want.site_cnt)
    assert np.array_equal(got.unit_rec_off, want.unit_rec_off) and got.n_candidates == pb.n_candidates
    # every unit's table rows came from the shard that owns the unit, in the shard's row order
    for u in range(pb.n_units):
        c = got.unit_counts(u)
        assert len(c) == len(got.unit_records(u))
        if len(c):
            assert (c[:, 0] % 10 == bin_of[u]).all() and (np.diff(c[:, 0] // 10) == 1).all()
