#!/usr/bin/env python
"""What the host side of an N-GPU box gives the pipelined step: plain cudaMemcpyAsync of the cfg2 step's byte
counts (57 MB in, 54 MB out per rank, pinned buffers, both directions at once), first one rank at a time
with the others idle, then all ranks together.  No kernels, nothing of this library: the denominator
for bench.py's end-to-end weak scaling.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py"""
import json
import os
import time

import torch
import torch.distributed as dist

N_IN, N_OUT, REPS = 57_000_000, 54_176_336, 20


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_in = torch.empty(N_IN, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(N_OUT, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(N_IN, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(N_OUT, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(active):
        barrier()
        if active:
            both()
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(REPS):
                both()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / REPS
        barrier()
        return dt

    alone = []
    for r in range(world):                                   # one rank copies, the others wait
        dt = timed(rank == r)
        t = torch.tensor([dt if rank == r else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        alone.append(float(t.item()))
    dt = timed(True)                                         # everybody copies
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    together = float(t.item())
    if rank == 0:
        per_rank = (N_IN + N_OUT) / 1e9
        print(json.dumps({
            "n_gpus": world, "bytes_in_per_rank": N_IN, "bytes_out_per_rank": N_OUT,
            "alone_ms_per_rank": [round(1e3 * a, 3) for a in alone],
            "alone_gb_s": [round(per_rank / a, 1) for a in alone],
            "together_ms": round(1e3 * together, 3),
            "together_aggregate_gb_s": round(world * per_rank / together, 1),
            "slowdown_together_vs_alone": round(together / (sum(alone) / len(alone)), 2),
            "what": "cudaMemcpyAsync H2D + D2H at once from pinned memory, max over ranks; no kernels"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
