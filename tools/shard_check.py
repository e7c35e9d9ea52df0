#!/usr/bin/env python
"""Sharded MI step on real GPUs, one process per GPU, checked against one GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/shard_check.py [--units 400]

Every rank draws the same heavy-tailed unit list, keeps the units lgmi_partition_lpt assigns to it,
runs them through the host API (api.mi_step_batched: host buffers in and out), and rank 0 gathers
the shards' results over torch.distributed (nccl) and restores the reference's row order.  Rank 0
then computes the whole list on its own GPU and compares rows, 3x3 tables' owner order, per-site
means and counts bit for bit.  Prints one JSON line; exit code 1 on any difference."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=400)
    ap.add_argument("--seed", type=int, default=20261033)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lg = importlib.import_module("l-giremi_b200")
    synth = importlib.import_module("l-giremi_b200.synth")
    shard = importlib.import_module("l-giremi_b200.shard")
    pb, _ = synth.make_heavy_tail(args.seed, args.units, s_max=300, r_max=3000)
    ctx = lg.Context(local)
    ok, info = True, {}
    for mode in (lg.MODE_HET_ONLY, lg.MODE_ALL_PAIRS):
        mine, index = shard.local_shard(pb, rank, world)
        res = lg.mi_step_batched(mine, 6, mode, ctx=ctx)
        merged = shard.gather_to_rank0(pb, res, index, device="cuda") if world > 1 else res
        if rank == 0:
            single = lg.mi_step_batched(pb, 6, mode, ctx=ctx)
            same = (np.array_equal(merged.records, single.records)
                    and np.array_equal(merged.site_mean, single.site_mean, equal_nan=True)
                    and np.array_equal(merged.site_cnt, single.site_cnt)
                    and np.array_equal(np.asarray(merged.unit_rec_off), np.asarray(single.unit_rec_off)))
            ok &= bool(same)
            info["mode_%d" % mode] = {"rows": int(single.n_records), "identical": bool(same)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"n_gpus": world, "units": args.units, "pairs": pb.n_candidates, "ok": ok, **info}), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
