#!/usr/bin/env python
"""BASELINE.json configs[3]: the heavy-tailed whole-transcriptome MI workload sharded over
N GPUs by cost-balanced (longest-processing-time) partitioning of the units.

    python tools/bench_cfg4.py [--units 20000] [--steps 10]                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/bench_cfg4.py                                    # N GPUs

STRONG scaling: the job (20 000 units, S ~ lognormal(ln 30, 0.8) in [2, 1000], R ~
lognormal(ln 150, 1.0) in [6, 20000]) is the same for every N; every rank draws the same batch
from the seed, keeps the units lgmi_partition_lpt assigns to it (cost S(S-1)/2 * ceil(R/64)),
uploads them once and runs the device-resident MI step.  No collective on the data path; the
barrier and the max-over-ranks time are the only uses of torch.distributed.  Rank 0 prints one
JSON line: whole-job site-pairs/s from the slowest rank's CUDA-event time, every rank's time
and modelled load (how good the cost model is as a balance criterion).  Not the headline bench
line (that is bench.py on cfg2); a parity-size check of the sharded result against a single
process is tests/test_shard_gloo.py."""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    # exactly ONE line on stdout: libraries that print banners to fd 1 (NCCL's version line) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--units", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seed", type=int, default=20261023)
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

    t0 = time.time()
    pb, _ = synth.make_heavy_tail(args.seed, args.units)
    gen_s = time.time() - t0
    S = pb.units["n_sites"].astype(np.int64)
    pairs_total = int((S * (S - 1) // 2).sum())
    bin_of, load = shard.plan(pb.units, world)
    index = np.nonzero(bin_of == rank)[0]
    mine = pb.subset(index)

    ctx = lg.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    batch = lg.Batch(ctx, mine)
    batch.upload()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    mc = 6
    for _ in range(max(3, args.warmup)):
        batch.run(mc, lg.MODE_ALL_PAIRS)
    res = batch.sync()
    n_records = int(res.n_records)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        flush.zero_()                                        # L2 flush between timed iterations (untimed)
        ev[k][0].record(stream)
        batch.run(mc, lg.MODE_ALL_PAIRS)
        ev[k][1].record(stream)
        batch.sync()
    barrier()
    my_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    stats = torch.tensor([my_ms, float(n_records), float(len(index)), float(mine.n_candidates)],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        every = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(every, stats)
    else:
        every = [stats]
    every = [e.cpu().tolist() for e in every]
    if rank == 0:
        ms = [e[0] for e in every]
        line = {
            "workload": "cfg4: %d heavy-tailed units (S~lognormal(ln30,0.8) in [2,1000], R~lognormal(ln150,1.0) "
                        "in [6,20000]), cov 0.5, mi_min_common_read %d, ALL_PAIRS" % (args.units, mc),
            "n_gpus": world, "scaling": "strong", "steps": args.steps, "pairs_per_step": pairs_total,
            "surviving_pairs_per_step": int(sum(e[1] for e in every)),
            "ms_per_step": max(ms), "value": pairs_total / (max(ms) * 1e-3), "unit": "site-pairs/s",
            "rank_ms": ms, "rank_units": [int(e[2]) for e in every], "rank_pairs": [int(e[3]) for e in every],
            "lpt_load": [int(x) for x in load], "lpt_load_max_over_mean": float(load.max() / load.mean()),
            "time_max_over_mean": max(ms) / (sum(ms) / len(ms)),
            "partition": "lgmi_partition_lpt on S(S-1)/2 * ceil(R/64)", "collectives_on_data_path": 0,
            "l2": "512 MiB buffer written between timed iterations", "generation_s": round(gen_s, 1),
        }
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
