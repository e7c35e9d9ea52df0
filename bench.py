#!/usr/bin/env python
"""bench.py -- site-pairs MI/s of the MI step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU MI step

Workload (N=1): BASELINE.json configs[1] -- 20 000 units (footprint x strand),
50 candidate sites x 200 reads each, coverage 0.5, mi_min_common_read 6:
24.5 M candidate site pairs per step.  One step = one pass of the whole MI
step (pair MI for every candidate, min-common filter, het filter, per-site
mean) over that batch.  For N>1 each rank owns its own 20 000-unit shard of an
N-times larger transcriptome (units are independent: no collective on the data
path; torch.distributed is used only for the barrier and the max-over-ranks
time), so scaling is "weak".

Prints ONE JSON line (rank 0):
  value         inputs resident in HBM, CUDA events around K steps, max over ranks
  e2e           the pipelined step through the C ABI: pinned HOST planes in (packed two-plane
                form), pinned HOST rows (MI + (i, j) arrays) + per-site means out, H2D / kernels /
                D2H of four groups of units overlapped; measured as a stream of batches
                (lgmi_pipeline_begin_packed / _collect / _finish over several pipelines: four steps
                in flight, every step uploading its input and reading its rows back) with the
                one-call-at-a-time figure (lgmi_pipeline_step_packed, four groups) and the plain
                three-plane upload-run-download time reported next to it
  roofline      the dominant kernel (k_pairs_fast) against the measured HBM bandwidth,
                timed live with CUDA events on the launching stream
  cpu_baseline  the unmodified reference (baseline/_ref; the oracle's port if that is
                absent) on the host cores over a bounded sample of the same units (N=1)
  het_only      rank 0's device-resident step in HET_ONLY | SKIP_NONHET mode (only the pairs the
                reference keeps are evaluated): reported separately, never in `value`
  dense         N=1 only: BASELINE.json configs[2], one unit of 2 000 sites x 100 000
                reads through the int8 tcgen05 path, with a tensor roofline for k_gram_i8
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import math
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "site-pairs MI/sec"
UNIT = "site-pairs/s"
CFG = dict(units=20000, sites=50, reads=200, cov=0.5, min_common=6, seed=20261020)
CFG3 = dict(sites=2000, reads=100000, cov=0.6, min_common=6, seed=20261021)
CFG4 = dict(units=20000, min_common=6, seed=20261023)       # heavy-tailed shapes: synth.heavy_tail_shapes
CFG5 = dict(units=20000, sites=50, reads=200, min_common=(6, 10, 20, 50), cov=(0.01, 0.05, 0.1, 0.25, 0.5), seed=20261030)
CFG1 = dict(seed=20261019, n_genes=40, reads_per_gene=250, threads=4, min_common=6, threshold=0.05)
CHUNK = 500
L2_FLUSH_BYTES = 512 << 20


def workload_name():
    return ("cfg2: %(units)d units x %(sites)d sites x %(reads)d reads, cov %(cov)g, "
            "mi_min_common_read %(min_common)d" % CFG)


# --------------------------------------------------------------------------- host cores / CPU arm
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def sample_units(n_units):
    """The first `n_units` units of rank 0's workload, in the reference's dict form."""
    synth = importlib.import_module("l-giremi_b200.synth")
    sb = synth.make_uniform(CFG["seed"], CHUNK, CFG["sites"], CFG["reads"], CFG["cov"], chunk=CHUNK)
    return [sb.mismatches(g) for g in range(min(n_units, CHUNK))]


def cpu_step(pool, cores, units, min_common):
    """The reference's MI step over `units` on `cores` processes, chunked the way
    giremi.py:367-370 chunks footprints.  Returns (pairs, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port
    n = max(1, len(units) // cores // 2)
    chunks = [(units[k:k + n], min_common) for k in range(0, len(units), n)]
    worker = ref_port.reference_chunk if ref_port.reference_functions() else ref_port.port_chunk
    t0 = time.perf_counter()
    out = pool.map(worker, chunks)
    dt = time.perf_counter() - t0
    return sum(p for p, _ in out), dt


def cpu_kind():
    """("reference" | "port", description): the unmodified reference installed under baseline/_ref
    when it is there, else the oracle's port of it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port
    if ref_port.reference_functions():
        return "reference", ("baseline/_ref = the unmodified reference (giremi 0.2.4, pip-installed): its own "
                             "mismatch_pair_mutual_info + mean_mismatch_pair_mutual_info, run as mismatch.py:387-404 does")
    return "port", "oracle/ref_port.py = the reference's per-pair dict rebuild + sklearn mutual_info_score"


def make_pool(cores):
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port  # noqa: F401  (imported before the fork so that workers have sklearn loaded)
    return mp.get_context("fork").Pool(cores)


def cpu_baseline(target_seconds=15.0):
    cores = host_cores()
    units = sample_units(CHUNK)
    with make_pool(cores) as pool:
        pairs, dt = cpu_step(pool, cores, units[:2 * cores], CFG["min_common"])       # calibration + warm-up
        rate = pairs / dt
        n = int(max(2 * cores, min(CHUNK, target_seconds * rate / 1225)))
        pairs, dt = cpu_step(pool, cores, units[:n], CFG["min_common"])
    kind, what = cpu_kind()
    return {"value": pairs / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "first %d of %d units of the workload (%d candidate pairs, %.1f s wall, mp.Pool(%d)); %s"
                      % (n, CFG["units"], pairs, dt, cores, what)}


def run_reference(args):
    """--impl reference: the reference's CPU MI step on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    cores = host_cores()
    units = sample_units(CHUNK)
    budget = 100.0                                           # seconds for the whole run
    with make_pool(cores) as pool:
        pairs, dt = cpu_step(pool, cores, units[:2 * cores], CFG["min_common"])
        rate = pairs / dt
        per_step = budget / max(1, args.steps + args.warmup)
        n = int(max(cores, min(CHUNK, per_step * rate / 1225)))
        for _ in range(args.warmup):
            cpu_step(pool, cores, units[:n], CFG["min_common"])
        tot_pairs, tot_dt = 0, 0.0
        for _ in range(args.steps):
            p, d = cpu_step(pool, cores, units[:n], CFG["min_common"])
            tot_pairs += p
            tot_dt += d
    value = tot_pairs / tot_dt
    kind, what = cpu_kind()
    sample = ("each step = first %d of %d units of the workload (%d candidate pairs) through mp.Pool(%d); "
              "throughput in pairs/s is size-independent for equal-shape units; %s" % (n, CFG["units"], n * 1225, cores, what))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (numpy/sklearn)",
        "data": "synthetic", "config": {"workload": workload_name(), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        # the timed region is tens of milliseconds: widen the window to the warm-up just before it,
        # which runs the same kernels, so that the median has several samples under load
        inside = [r for t, r in self.rows if t0 - 1.0 <= t <= t1 + 0.05] or [r for _, r in self.rows[-3:]]
        sm, smax, reasons = [], [], set()
        for row in inside:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- deep unit (cfg3), tensor-core path
def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {}, "fallback (B200_PROFILING.md)"


def int8_library_tops(torch):
    """cuBLASLt int8 GEMM 8192^3 through torch._int_mm, best of 10: context for the int8 peak."""
    try:
        a = torch.randint(-4, 4, (8192, 8192), dtype=torch.int8, device="cuda")
        b = torch.randint(-4, 4, (8192, 8192), dtype=torch.int8, device="cuda").t()
        for _ in range(3):
            torch._int_mm(a, b)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception:                                        # noqa: BLE001 (context figure only)
        return None


def dense_leg(lg, synth, ctx, torch, stream, steps, warmup, flush):
    """BASELINE.json configs[2]: one unit of 2 000 sites x 100 000 reads through
    k_dense_prep + k_dense_x + k_gram_i8 (int8 tcgen05) + k_other_fix + the MI epilogue kernels."""
    S, R, mc = CFG3["sites"], CFG3["reads"], CFG3["min_common"]
    pb, _ = synth.make_deep_unit(CFG3["seed"], S, R, CFG3["cov"])
    batch = lg.Batch(ctx, pb)
    batch.upload()
    for _ in range(warmup):
        batch.run(mc, lg.MODE_ALL_PAIRS)
    res = batch.sync()
    assert int(res.n_dense_units) == 1, "cfg3 must take the tensor-core path"
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    gram_ms, dense_ms = [], []
    for k in range(steps):
        flush.zero_()
        ev[k][0].record(stream)
        batch.run(mc, lg.MODE_ALL_PAIRS)
        ev[k][1].record(stream)
        r = batch.sync()
        gram_ms.append(float(r.gram_kernel_ms))
        dense_ms.append(float(r.dense_kernel_ms))
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    pairs = S * (S - 1) // 2
    algo_ops = 2 * 9 * pairs * R                             # nine count matrices over the upper triangle
    g_ms = sum(gram_ms) / len(gram_ms)
    peaks, src = load_peaks()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_gram_i8_dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    bf16 = float(peaks.get("bf16_tflops", 1590.0))
    proxy = 2.0 * bf16                                       # int8 dense = 2x bf16 dense on sm_100
    lib = int8_library_tops(torch)                           # cuBLASLt int8 8192^3 measured on this box, now
    peak = max(lib, proxy) if lib else proxy                 # the measured int8 peak where the library reaches it
    out = {
        "workload": "cfg3: 1 unit x %d sites x %d reads, cov %g, mi_min_common_read %d" % (S, R, CFG3["cov"], mc),
        "pairs_per_step": pairs, "surviving_pairs": int(r.n_records), "ms_per_step": step_ms,
        "value": pairs / (step_ms * 1e-3), "unit": UNIT,
        "roofline": {"bound": "tensor", "achieved": algo_ops / (g_ms * 1e-3) / 1e12, "peak": peak, "unit": "TOP/s",
                     "frac": algo_ops / (g_ms * 1e-3) / 1e12 / peak, "traffic": traffic, "kernel": "k_gram_i8",
                     "kernel_ms": g_ms, "algorithmic_ops": algo_ops, "issued_ops": 2 * int(r.gram_macs),
                     "issued_tops": 2 * int(r.gram_macs) / (g_ms * 1e-3) / 1e12,
                     "peak_source": ("max(cuBLASLt int8 8192^3 measured live = %.0f, 2 x bf16_tflops burst = %.0f; %s)"
                                     % (lib, proxy, src)) if lib else "2 x bf16_tflops burst, %s" % src,
                     "library_int8_tops": lib, "peak_proxy_2x_bf16": proxy,
                     "frac_of_proxy": algo_ops / (g_ms * 1e-3) / 1e12 / proxy,
                     "issued_frac": 2 * int(r.gram_macs) / (g_ms * 1e-3) / 1e12 / peak,
                     "kernel_share_of_step": g_ms / step_ms,
                     "note": "achieved = the contract's algorithmic count (nine count matrices over the upper triangle) / "
                             "kernel time; in the four-block form the kernel issues 4/9 of that work (issued_ops, "
                             "issued_frac = the tensor pipe's own utilisation) and k_other_fix counts the 'other' cells "
                             "beside it on a second stream, which is inside kernel_ms"},
        "form": ("four Gram blocks (P, M x P, M) + 'other' cells from the listed reads" if int(r.n_dense_four)
                 else "nine Gram blocks (C, P, M x C, P, M)"),
        "prep_plus_gram_ms": sum(dense_ms) / len(dense_ms),
    }
    batch.close()
    return out


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_near_gpu(torch, local):
    """Pin this rank's threads to the CPUs next to its GPU's PCIe root, so that the pinned staging
    buffers (first touched afterwards) and the copy threads sit on that socket.  Sources, in order:
    sysfs numa_node of the GPU, sysfs local_cpulist, `nvidia-smi topo -m`'s CPU-affinity column.
    Returns the previous affinity (restored before the CPU baseline) and a note for the JSON line."""
    try:
        prev = os.sched_getaffinity(0)
    except AttributeError:
        return None, "not bound (no sched_getaffinity)"
    props = torch.cuda.get_device_properties(local)
    cpus, how = set(), None
    try:
        base = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open(base + "/numa_node").read().strip())
        if node >= 0:
            cpus = _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
            how = "numa node %d (sysfs)" % node
        else:
            cpus = _cpulist(open(base + "/local_cpulist").read())
            how = "local_cpulist (sysfs, numa_node=-1)"
    except (OSError, ValueError, AttributeError):
        pass
    if not cpus or cpus >= prev:            # nothing, or "every cpu": ask the driver
        try:
            out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            out = re.sub(r"\x1b\[[0-9;]*m", "", out)               # the header row is underlined with ANSI codes
            header = None
            for line in out.splitlines():
                cells = [c.strip() for c in line.split("\t") if c.strip() != ""]
                if header is None and "CPU Affinity" in cells:
                    header = cells
                elif header is not None and cells and cells[0] == "GPU%d" % local:
                    col = header.index("CPU Affinity") + 1          # the row carries the row label first
                    got = _cpulist(cells[col])
                    numa = cells[col + 1] if col + 1 < len(cells) else "?"
                    if got:
                        cpus, how = got, "numa node %s (nvidia-smi topo)" % numa
                    break
        except (OSError, ValueError, IndexError, subprocess.SubprocessError):
            pass
    cpus &= prev
    if not cpus:
        return prev, "not bound (no topology information for this GPU)"
    if cpus == prev:
        return prev, "all %d cpus are local: %s" % (len(prev), how)
    os.sched_setaffinity(0, cpus)
    return prev, "bound to %s, %d cpus" % (how, len(cpus))


# --------------------------------------------------------------------------- strong scaling (cfg4)
def result_hash(res, n_units):
    """sha256 over what the MI step hands on: rows (unit, i, j, mi bits) in order, per-site mean and count."""
    import numpy as np
    h = hashlib.sha256()
    rec = res.records
    h.update(np.ascontiguousarray(rec['unit']).tobytes())
    h.update(np.ascontiguousarray(rec['i']).tobytes())
    h.update(np.ascontiguousarray(rec['j']).tobytes())
    h.update(np.ascontiguousarray(rec['mi']).view(np.uint64).tobytes())
    mean = np.ascontiguousarray(res.site_mean, dtype=np.float64).copy()
    mean[np.isnan(mean)] = np.nan                            # one NaN pattern
    h.update(mean.view(np.uint64).tobytes())
    h.update(np.ascontiguousarray(res.site_cnt, dtype=np.uint32).tobytes())
    h.update(np.ascontiguousarray(res.unit_rec_off, dtype=np.uint64)[:n_units + 1].tobytes())
    return h.hexdigest()


def strong_leg(lg, synth, ctx, torch, dist, rank, world, stream, steps, warmup, flush, barrier):
    """BASELINE.json configs[3]: ONE global list of heavy-tailed units, partitioned over the ranks by
    lgmi_partition_lpt on S(S-1)/2 * ceil(R/64).  Each rank takes its shard through the host API
    (pinned host planes in, pinned host rows + means out: lgmi_pipeline_step_packed), rank 0 gathers
    the shards' results (torch.distributed, no collective on the data path) and restores the
    reference's row order; the merged result must hash to what ONE GPU computes for the whole list."""
    import numpy as np
    shard = importlib.import_module("l-giremi_b200.shard")
    mc = CFG4["min_common"]
    mode = lg.MODE_HET_ONLY                                  # every candidate evaluated, het-kept rows returned
    t0 = time.perf_counter()
    pb_all, _ = synth.make_heavy_tail(CFG4["seed"], CFG4["units"])
    gen_s = time.perf_counter() - t0
    costs = lg.unit_costs(pb_all.units)
    bin_of, load = lg.partition_lpt(costs, world)
    index = np.flatnonzero(bin_of == rank)
    mine = pb_all.subset(index) if world > 1 else pb_all
    pairs_total = pb_all.n_candidates

    def host_api(pb):
        """(step function, pipeline, pinned buffers) of the host-buffer step over `pb`."""
        chunks = 4 if pb.n_candidates >= 8_000_000 else (2 if pb.n_candidates >= 2_000_000 else 1)
        pipe = lg.Pipeline(ctx, pb, chunks)
        packed = pb.packed2(tight=True)
        pin_p = ctx.pinned_empty(packed.shape, np.uint32)
        pin_f = ctx.pinned_empty(pb.site_flags.shape, np.uint8)
        pin_p.array[...] = packed
        pin_f.array[...] = pb.site_flags
        fn = lambda copy=False: pipe.step(mc, mode | lg.MODE_COMPACT_OUTPUT, pin_p.array, pin_f.array, copy=copy, tight=True)
        return fn, pipe, (pin_p, pin_f), chunks, packed.nbytes + pb.site_flags.nbytes

    def timed_host(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t) / steps

    def timed_host_stream(pb, pins, depth=3):
        """The same host-buffer step as a stream of batches: `depth` steps in flight over as many one-group
        pipelines (lgmi_pipeline_begin_packed / _collect / _finish); starts idle, ends drained."""
        pipes = [lg.Pipeline(ctx, pb, 1) for _ in range(depth)]
        def run(n):
            feed = ((pins[0].array, pins[1].array) for _ in range(n))
            for _k, _res in lg.stream_steps(pipes, feed, mc, mode | lg.MODE_COMPACT_OUTPUT, copy=False, tight=True):
                pass
        run(max(depth, warmup))
        torch.cuda.synchronize()
        t = time.perf_counter()
        run(steps)
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t) / steps
        for q in pipes:
            q.close()
        return ms

    def timed_device(pb):
        batch = lg.Batch(ctx, pb)
        batch.upload()
        mode_dev = lg.MODE_ALL_PAIRS | lg.MODE_GRAPH         # (as the headline leg: the launch chain replayed as one CUDA graph)
        for _ in range(warmup):
            batch.run(mc, mode_dev)
        batch.sync()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.zero_()
            ev[k][0].record(stream)
            batch.run(mc, mode_dev)
            ev[k][1].record(stream)
            batch.sync()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        batch.close()
        return ms

    def all_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world == 1:
            return [float(x)]
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        return [float(e.item()) for e in every]

    step, pipe, pins, chunks, h2d = host_api(mine)
    barrier()
    host_ms = all_ranks(timed_host(step))                    # upload + kernels + download of this rank's shard
    barrier()
    stream_ms = all_ranks(timed_host_stream(mine, pins))     # ... with three steps in flight
    barrier()
    dev_ms = all_ranks(timed_device(mine))                   # the same shard resident in HBM, ALL_PAIRS, CUDA events
    barrier()
    res = step(copy=True)
    t0 = time.perf_counter()
    merged = shard.gather_to_rank0(pb_all, res, index, device="cuda") if world > 1 else res
    gather_ms = 1e3 * (time.perf_counter() - t0)
    d2h = int(res.n_records) * (8 + res.rec_ij.dtype.itemsize) + mine.n_sites * 8 + (mine.n_units + 1) * 8
    pipe.close()
    out = None
    if rank == 0:
        sharded_hash = result_hash(merged, pb_all.n_units)
        # the whole list on this one GPU: the reference result and the N = 1 time of this very run
        if world > 1:
            step1, pipe1, pins1, _c, _b = host_api(pb_all)
            single = step1(copy=True)
            single_hash = result_hash(single, pb_all.n_units)
            t1_host = timed_host(step1)
            pipe1.close()
            t1_dev = timed_device(pb_all)
        else:
            b = lg.Batch(ctx, pb_all)                        # N = 1: the plain upload / run / download path as the check
            b.upload()
            b.run(mc, mode)
            single = b.download(copy=True)
            single_hash = result_hash(single, pb_all.n_units)
            b.close()
            t1_host, t1_dev = host_ms[0], dev_ms[0]
        if sharded_hash != single_hash:
            raise SystemExit("bench.py: the result gathered from %d shards differs from the single-GPU result" % world)
        out = {
            "workload": "cfg4: %d heavy-tailed units (S~lognormal(ln30,0.8) in [2,1000], R~lognormal(ln150,1.0) in "
                        "[6,20000]), cov 0.5, mi_min_common_read %d; ONE global list for every N" % (CFG4["units"], mc),
            "scaling": "strong", "n_gpus": world, "pairs_per_step": pairs_total,
            "partition": "lgmi_partition_lpt on S(S-1)/2 * ceil(R/64); collectives on the data path: 0",
            "lpt_load_max_over_mean": float(load.max() / load.mean()),
            "ms_per_step": max(host_ms), "value": pairs_total / (max(host_ms) * 1e-3), "unit": UNIT,
            "what": "per rank: lgmi_pipeline_step_packed over pinned host buffers (H2D + kernels + D2H, HET_ONLY rows + "
                    "per-site means; tight two-plane input, compact rows), wall clock, max over ranks",
            "rank_ms": host_ms, "n1_ms_per_step": t1_host, "efficiency_vs_n1": t1_host / (world * max(host_ms)),
            "streamed": {"what": "the same host-buffer step as a stream of batches: three steps in flight per rank "
                                 "(lgmi_pipeline_begin_packed / _collect / _finish, one group per step), wall clock over "
                                 "the steps from idle to drained, max over ranks",
                         "ms_per_step": max(stream_ms), "value": pairs_total / (max(stream_ms) * 1e-3), "rank_ms": stream_ms},
            "device_resident": {"what": "the shard resident in HBM, ALL_PAIRS, launch chain as one CUDA graph, CUDA events, max over ranks",
                                "ms_per_step": max(dev_ms), "value": pairs_total / (max(dev_ms) * 1e-3),
                                "rank_ms": dev_ms, "n1_ms_per_step": t1_dev,
                                "efficiency_vs_n1": t1_dev / (world * max(dev_ms)),
                                "time_max_over_mean": max(dev_ms) / (sum(dev_ms) / len(dev_ms))},
            "gather_to_rank0_ms": gather_ms, "rows": int(merged.n_records),
            "records_sha256": sharded_hash, "single_gpu_sha256": single_hash, "identical_to_single_gpu": True,
            "h2d_bytes_rank0": h2d, "d2h_bytes_rank0": d2h, "pipeline_groups_rank0": chunks, "generation_s": round(gen_s, 1),
        }
    barrier()
    return out


# --------------------------------------------------------------------------- sparsity sweep (cfg5)
def cfg5_leg(lg, synth, ctx, torch, dist, rank, world, stream, flush, peak_gbs):
    """BASELINE.json configs[4]: mi_min_common_read x coverage grid, 20 000 units of 50 sites x 200 reads
    per point, split evenly over the ranks (equal shapes: the LPT partition is the even split).  Device-resident
    ALL_PAIRS step per point: CUDA events, max over ranks; per-point roofline of rank 0's k_pairs_fast."""
    import numpy as np
    G = CFG5["units"] // world
    points = []
    for ci, cov in enumerate(CFG5["cov"]):
        pb = synth.make_uniform_planes(CFG5["seed"] + 100 * ci + rank, G, CFG5["sites"], CFG5["reads"], cov, chunk=CHUNK)
        batch = lg.Batch(ctx, pb)
        batch.upload()
        for mc in CFG5["min_common"]:
            for _ in range(3):
                batch.run(mc, lg.MODE_ALL_PAIRS)
            batch.sync()
            n = 5
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
            k_ms = []
            for k in range(n):
                flush.zero_()
                ev[k][0].record(stream)
                batch.run(mc, lg.MODE_ALL_PAIRS)
                ev[k][1].record(stream)
                r = batch.sync()
                k_ms.append(float(r.pairs_kernel_ms))
            ms = sum(a.elapsed_time(b) for a, b in ev) / n
            t = torch.tensor([ms, float(r.n_records)], dtype=torch.float64, device="cuda")
            tmax, tsum = t.clone(), t.clone()
            if world > 1:
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            kms = sum(k_ms) / len(k_ms)
            ab = batch.algorithmic_bytes()
            points.append({"cov": cov, "min_common": mc, "ms_per_step": float(tmax[0].item()),
                           "value": pb.n_candidates * world / (float(tmax[0].item()) * 1e-3),
                           "surviving_fraction": float(tsum[1].item()) / (pb.n_candidates * world),
                           "roofline": {"bound": "hbm", "kernel": "k_pairs_fast", "kernel_ms": kms,
                                        "achieved": ab / (kms * 1e-3) / 1e9 if kms > 0 else None, "peak": peak_gbs,
                                        "unit": "GB/s", "frac": ab / (kms * 1e-3) / 1e9 / peak_gbs if kms > 0 else None,
                                        "algorithmic_bytes": ab}})
        batch.close()
    if rank != 0:
        return None
    return {"workload": "cfg5: %d units x %d sites x %d reads per grid point over %d GPU(s), device-resident ALL_PAIRS"
                        % (G * world, CFG5["sites"], CFG5["reads"], world),
            "unit": UNIT, "points": points,
            "parity": "tests/test_gpu_parity.py::test_cfg5_grid_against_oracle (oracle size) and "
                      "::test_cfg5_full_size_oracle_sample (every grid point at full size)"}


# --------------------------------------------------------------------------- the CLI (cfg1)
def cfg1_leg(n_gpus):
    """BASELINE.json configs[0]: the reference's own `l-giremi` CLI (unmodified, baseline/_ref; pysam replaced by
    the stand-in over a simulated single-chromosome dataset) run stock on the host cores and with this
    repository's MI step patched in (install(batched=True): workers extract, the parent owns the GPUs).
    Wall time of main() and of the MI step inside it; the tables must agree."""
    import pickle
    import tempfile
    import numpy as np
    import pandas as pd
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import simdata
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "giremi")):
        return {"unavailable": "baseline/_ref (the installed reference) is not in this checkout"}
    threads = max(CFG1["threads"], n_gpus)
    with tempfile.TemporaryDirectory() as tmp:
        ds = simdata.Dataset(seed=CFG1["seed"], n_genes=CFG1["n_genes"], reads_per_gene=CFG1["reads_per_gene"])
        with open(os.path.join(tmp, "ds.pkl"), "wb") as fh:
            pickle.dump(ds, fh)
        extra = ["-t", str(threads), "--mi_min_common_read", str(CFG1["min_common"]), "--mi_p_threshold",
                 str(CFG1["threshold"])]
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LGMI_DEVICE")}
        env["LGMI_DEVICES"] = ",".join(str(d) for d in range(n_gpus))
        runs = {}
        for name, flag in (("stock", []), ("patched", ["--patched"])):
            cmd = [sys.executable, os.path.join(ROOT, "tools", "run_cli.py")] + flag + \
                  [os.path.join(tmp, "ds.pkl"), os.path.join(tmp, name)] + extra
            r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1200)
            if r.returncode != 0:
                return {"error": "%s CLI run failed: %s" % (name, r.stderr[-400:])}
            runs[name] = json.loads(r.stdout.strip().splitlines()[-1])
        same = True
        rows = {}
        for ext, cols in ((".mi.txt", ("mi",)), (".mismatch.txt", ("mean_mi", "mip"))):
            a, b = pd.read_table(os.path.join(tmp, "stock" + ext)), pd.read_table(os.path.join(tmp, "patched" + ext))
            rows[ext] = len(a)
            same &= len(a) == len(b) and list(a.columns) == list(b.columns)
            for c in a.columns:
                if not same:
                    break
                if c == "score":                             # model trained on an unseeded sample (giremi.py:117-122)
                    continue
                if c in cols or a[c].dtype.kind == "f":
                    x, y = a[c].to_numpy(dtype=float), b[c].to_numpy(dtype=float)
                    same &= bool(np.array_equal(np.isnan(x), np.isnan(y)) and
                                 np.all(np.isnan(y) | (np.abs(x - y) <= 1e-10 * np.abs(y) + 1e-15)))
                else:
                    same &= bool(a[c].equals(b[c]))
    st, pa = runs["stock"], runs["patched"]
    return {
        "workload": "cfg1: l-giremi -t %d on a simulated single-chromosome dataset (%d genes, %d spliced reads with cs "
                    "tags, pysam stand-in)" % (threads, CFG1["n_genes"], CFG1["n_genes"] * CFG1["reads_per_gene"]),
        "stock": {"wall_s": st["wall_s"], "mi_step_cpu_s": st["mi_step_cpu_s"], "mi_units": st["mi_calls"],
                  "what": "the two MI functions timed inside the pool workers, seconds summed over workers"},
        "patched": {"wall_s": pa["wall_s"], "mi_step_s": pa.get("mi_step_s"), "gpu_submit_s": pa.get("gpu_submit_s"),
                    "mip_s": pa.get("mip_s"), "extract_pool_s": pa.get("extract_pool_s"),
                    "cuda_start_s": pa.get("cuda_start_s"), "n_gpus": pa.get("n_gpus"), "mi_units": pa.get("n_units"),
                    "what": "workers extract + encode (extract_pool_s) while the parent starts CUDA (cuda_start_s, "
                            "hidden behind the extraction); mi_step_s = the parent's MI step over all units (concatenate, "
                            "submit, build the frames), gpu_submit_s = the submit alone; mip = one lgmi_ecdf launch"},
        "tables_identical": bool(same), "rows": rows,
        "mi_step_speedup": st["mi_step_cpu_s"] / pa["mi_step_s"] if pa.get("mi_step_s") else None,
    }


# --------------------------------------------------------------------------- GPU arm
def link_probe(torch, dist, rank, world, h2d_bytes, d2h_bytes, reps=8):
    """What the host side of the box gives this step: plain cudaMemcpyAsync of the step's byte counts from / to
    pinned memory, both directions at once, no kernels -- this rank alone (the others idle) and all ranks together.
    The end-to-end leg cannot be faster than `together_ms`; on a box whose GPUs share the host fabric that, not
    the kernels, is what the end-to-end weak scaling shows."""
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(active):
        sync_all()
        if active:
            both()
        sync_all()
        t0 = time.perf_counter()
        if active:
            for _ in range(reps):
                both()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        sync_all()
        return dt

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    alone = max_over_ranks(timed(rank == 0) if world > 1 else timed(True))     # rank 0 copies, the others wait
    together = max_over_ranks(timed(True)) if world > 1 else alone
    return {"what": "cudaMemcpyAsync of this step's bytes from / to pinned memory, both directions at once, no kernels; "
                    "rank 0 alone, then all ranks together (max over ranks)",
            "alone_ms": 1e3 * alone, "together_ms": 1e3 * together,
            "together_aggregate_gb_s": world * (h2d_bytes + d2h_bytes) / together / 1e9}


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the MI step has no CPU path (use --impl reference for the CPU arm)")
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    prev_affinity, numa_note = bind_near_gpu(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lg = importlib.import_module("l-giremi_b200")
    synth = importlib.import_module("l-giremi_b200.synth")
    ctx = lg.Context(local)
    stream = torch.cuda.Stream()                             # a real (non-default) stream: handle 0 would mean
    torch.cuda.set_stream(stream)                            # "library's own stream" to lgmi_set_stream
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)                       # torch.cuda.Event sees the library's launches
    if args.dense_only:
        flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
        print(json.dumps({"dense": dense_leg(lg, synth, ctx, torch, stream, args.steps, args.warmup, flush)}), flush=True)
        return 0

    # this rank's shard: LPT over the global unit list (equal costs -> equal bins); its data by seed
    G = CFG["units"]
    costs = np.full(G * world, lg.unit_costs(np.array([(0, CFG["sites"], CFG["reads"], 8, 0)],
                                                      dtype=lg.UNIT_DESC))[0], dtype=np.uint64)
    bin_of, load = lg.partition_lpt(costs, world)
    assert int((bin_of == rank).sum()) == G and load.max() == load.min()
    pb = synth.make_uniform_planes(CFG["seed"] + rank, G, CFG["sites"], CFG["reads"], CFG["cov"], chunk=CHUNK)
    pairs_per_step = pb.n_candidates

    # pinned host staging (the public API's input buffers)
    pin_planes = ctx.pinned_empty(pb.planes.shape, np.uint32)
    pin_flags = ctx.pinned_empty(pb.site_flags.shape, np.uint8)
    pin_planes.array[...] = pb.planes
    pin_flags.array[...] = pb.site_flags
    batch = lg.Batch(ctx, pb)
    batch.upload(pin_planes.array, pin_flags.array)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    mode_dev = lg.MODE_ALL_PAIRS
    mc = CFG["min_common"]

    # ---- device-resident throughput ------------------------------------------------------------
    for _ in range(args.warmup):
        batch.run(mc, mode_dev)
    res = batch.sync()
    n_records = int(res.n_records)
    algo_bytes = batch.algorithmic_bytes()
    # the dominant kernel alone, timed live with CUDA events on the launching stream (the library brackets
    # k_pairs_fast with its own events when the chain is launched kernel by kernel)
    pairs_ms = []
    for _ in range(max(3, min(args.steps, 10))):
        flush.zero_()
        batch.run(mc, mode_dev)
        pairs_ms.append(float(batch.sync().pairs_kernel_ms))
    # the timed steps replay the same launch chain as one CUDA graph: no host launch gaps between the kernels
    mode_timed = mode_dev | lg.MODE_GRAPH
    for _ in range(args.warmup):
        batch.run(mc, mode_timed)
    batch.sync()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 1.0:                # ~1 s of the same kernels under the clock sampler
        batch.run(mc, mode_timed)
        batch.sync()
    launches0 = ctx.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                        # L2 flush between timed iterations (untimed)
        ev[k][0].record(stream)
        batch.run(mc, mode_timed)
        ev[k][1].record(stream)
        r = batch.sync()
        assert int(r.n_records) == n_records
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    launches = ctx.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = pairs_per_step * world * args.steps / (dev_ms_max * 1e-3)

    # ---- the same step restricted to what the reference keeps (reported separately) ----
    # HET_ONLY | SKIP_NONHET: only pairs next to a het SNP are evaluated (mismatch.py:393-396 drops the
    # others right after computing them); rows and per-site means are identical to the HET_ONLY output.
    mode_het = lg.MODE_HET_ONLY | lg.MODE_SKIP_NONHET
    for _ in range(3):
        batch.run(mc, mode_het)
    rh = batch.sync()
    het_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        het_ev[k][0].record(stream)
        batch.run(mc, mode_het)
        het_ev[k][1].record(stream)
        batch.sync()
    het_ms = sum(a.elapsed_time(b) for a, b in het_ev) / args.steps
    is_het = (pb.site_flags & 3) == lg.SITE_HET_SNP
    n_het = np.add.reduceat(is_het.astype(np.int64), pb.units["site_off"].astype(np.int64))
    n_s = pb.units["n_sites"].astype(np.int64)
    het_pairs = int((n_s * (n_s - 1) // 2 - (n_s - n_het) * (n_s - n_het - 1) // 2).sum())
    het_leg = {"mode": "HET_ONLY | SKIP_NONHET: only pairs next to a het SNP evaluated (same rows and means as HET_ONLY)",
               "ms_per_step": het_ms, "evaluated_pairs_per_step": het_pairs, "rows_per_step": int(rh.n_records),
               "candidate_pairs_per_s": pairs_per_step / (het_ms * 1e-3), "evaluated_pairs_per_s": het_pairs / (het_ms * 1e-3)}

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    # lgmi_pipeline_step: pinned host planes in, pinned host rows + per-site means out; H2D, kernels and
    # D2H of consecutive groups of units overlap inside the call (what lg.mi_step_batched does for big batches)
    mode_e2e = lg.MODE_HET_ONLY                              # what mismatch.py:393-404 hands on
    pipe = lg.Pipeline(ctx, pb, args.e2e_chunks)
    # the library's compact host formats: two-plane input without the 128-read padding (2 bits per site and read),
    # rows as an MI array + a 2-byte (i, j) array, no per-site count (it is the number of rows a site appears in)
    packed = pb.packed2(tight=True)
    pin_packed = ctx.pinned_empty(packed.shape, np.uint32)
    pin_packed.array[...] = packed
    def e2e_step():
        return pipe.step(mc, mode_e2e | lg.MODE_COMPACT_OUTPUT, pin_packed.array, pin_flags.array, copy=False, tight=True)
    # the same step as a stream of batches: a second pipeline, and step k + 1 begun (lgmi_pipeline_begin_packed)
    # before step k is collected (lgmi_pipeline_finish) -- every step still uploads its input and reads its rows
    depth = max(2, args.e2e_depth)                           # steps in flight = pipelines
    # (with several steps in flight a step need not be cut into groups: one group per step by default)
    more = [] if args.no_e2e else [lg.Pipeline(ctx, pb, args.e2e_stream_chunks) for _ in range(depth)]
    def e2e_stream(n):                                       # begin(k); collect(k - depth + 2); finish(k - depth + 1)
        res = None
        feed = ((pin_packed.array, pin_flags.array) for _ in range(n))
        for _k, res in lg.stream_steps(more, feed, mc, mode_e2e | lg.MODE_COMPACT_OUTPUT, copy=False, tight=True,
                                       collect=bool(args.e2e_collect)):
            pass
        return res
    e2e_steps = 0 if args.no_e2e else args.steps
    for _ in range(args.warmup if e2e_steps else 1):
        out = e2e_step()
    h2d = packed.nbytes + pb.site_flags.nbytes
    groups_e2e = args.e2e_stream_chunks                      # (of the streamed measurement, which `value` reports)
    d2h = out.n_records * (8 + out.rec_ij.dtype.itemsize) + pb.n_sites * 8 + (pb.n_units + groups_e2e) * 8 + 16 * groups_e2e
    launches_e2e0 = ctx.launch_count
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    launches_e2e = ctx.launch_count - launches_e2e0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    barrier()
    e2e_sync_ms = 1e3 * float(t.item()) / e2e_steps if e2e_steps else None
    e2e_sync_value = pairs_per_step * world * e2e_steps / float(t.item()) if e2e_steps else None
    e2e_records = int(out.n_records)
    e2e_value, e2e_ms = e2e_sync_value, e2e_sync_ms
    if e2e_steps:
        outs = e2e_stream(max(depth, args.warmup))
        assert outs.n_records == e2e_records
        launches_e2e0 = ctx.launch_count
        barrier()
        t0 = time.perf_counter()
        outs = e2e_stream(e2e_steps)                          # starts idle, ends drained: fill and drain are inside
        torch.cuda.synchronize()
        stream_s = time.perf_counter() - t0
        launches_e2e = ctx.launch_count - launches_e2e0
        t = torch.tensor([stream_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        assert outs.n_records == e2e_records
        e2e_ms = 1e3 * float(t.item()) / e2e_steps
        e2e_value = pairs_per_step * world * e2e_steps / float(t.item())
        for q in more:
            q.close()
    link = link_probe(torch, dist, rank, world, int(h2d), int(d2h)) if e2e_steps else None
    # the plain sequence (upload, run, download one after the other), for comparison
    def serial_step():
        batch.upload(pin_planes.array, pin_flags.array)
        batch.run(mc, mode_e2e)
        return batch.download(copy=False)
    serial_ms = None
    if e2e_steps:
        serial_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            serial_step()
        torch.cuda.synchronize()
        serial_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps

    peaks = {}
    peak_src = "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- the other configs of BASELINE.json, every rank taking part -----------------------------
    batch.close()
    pipe.close()
    strong = None if args.no_strong else strong_leg(lg, synth, ctx, torch, dist, rank, world, stream,
                                                    max(3, min(args.steps, 10)), 3, flush, barrier)
    cfg5 = None if args.no_cfg5 else cfg5_leg(lg, synth, ctx, torch, dist, rank, world, stream, flush, peak)
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    k_ms = sum(pairs_ms) / len(pairs_ms)
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("k_pairs_fast_dram_bytes_per_launch")
        except (OSError, ValueError):
            traffic = None

    if prev_affinity:
        os.sched_setaffinity(0, prev_affinity)               # the CPU baseline gets every host core again
    cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline()      # rank 0 at N=1 only
    dense = None
    if world == 1 and not args.no_dense:
        dense = dense_leg(lg, synth, ctx, torch, stream, max(3, min(args.steps, 5)), 3, flush)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 popcount + f64 MI", "data": "synthetic",
        "config": {"workload": workload_name(), "units_per_gpu": G, "pairs_per_step_per_gpu": pairs_per_step,
                   "surviving_pairs_per_step_per_gpu": n_records, "mode": "ALL_PAIRS (every candidate evaluated and "
                   "every survivor written)", "l2": "512 MiB buffer written between timed iterations",
                   "launch": "the step's kernel chain replayed as one CUDA graph (LGMI_MODE_GRAPH)",
                   "partition": "LPT over %d units -> %d bins, loads equal" % (G * world, world)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_pairs_fast", "kernel_ms": k_ms, "algorithmic_bytes": algo_bytes,
                     "peak_source": peak_src,
                     "kernel_share_of_step": k_ms * args.steps / dev_ms if dev_ms else None,
                     "note": "HBM is the contract's roofline for this path, not what bounds it: the bit-exact fp64 epilogue "
                             "(about 135 fp64 instructions per pair at 64 lanes/clk/SM) alone caps the kernel near 0.41 of "
                             "this peak, and AND/popcount + bookkeeping share the issue slots with it (DESIGN.md section 5)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "mode": "HET_ONLY: all candidates evaluated, het-kept rows (%d) + per-site mean MI returned"
                        % e2e_records, "ms_per_step": e2e_ms,
                "api": "lgmi_pipeline_begin_packed / _collect / _finish on %d pipelines, step k + %d begun before step k "
                       "is finished (a stream of batches; every step uploads its input from pinned host memory and reads "
                       "its rows back; the timed region starts idle and ends drained); tight two-plane input, MI + 2-byte "
                       "(i, j) rows, no per-site count, %d group(s) of units per step, each group's kernels one "
                       "CUDA graph" % (depth, depth - 1, args.e2e_stream_chunks),
                "steps_in_flight": depth,
                "one_step_at_a_time": {"ms_per_step": e2e_sync_ms, "value": e2e_sync_value, "unit": UNIT,
                                       "api": "lgmi_pipeline_step_packed, each call returning before the next starts; "
                                              "%d groups of units on their own streams" % args.e2e_chunks},
                "serial_upload_run_download_ms": serial_ms, "link_probe": link, "gpu_launches": launches_e2e * world,
                "host_affinity": numa_note},
        "gpu_launches": launches * world,
        "clocks": clocks,
    }
    line["het_only"] = het_leg
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if dense is not None:
        line["dense"] = dense
    if strong is not None:
        line["strong"] = strong
    if cfg5 is not None:
        line["cfg5"] = cfg5
    if not args.no_cfg1:
        line["cfg1"] = cfg1_leg(world)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # exactly ONE line on stdout: libraries that print banners to fd 1 (NCCL's version line) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["graft", "reference"], default="graft")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end leg")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="groups of units of the pipelined end-to-end step")
    ap.add_argument("--e2e-depth", type=int, default=4, help="steps in flight of the streamed end-to-end measurement")
    ap.add_argument("--e2e-stream-chunks", type=int, default=1, help="groups of units per step of the streamed measurement")
    ap.add_argument("--e2e-collect", type=int, default=1, help="0: no lgmi_pipeline_collect ahead of the finish")
    ap.add_argument("--no-dense", action="store_true", help="skip the cfg3 deep-unit (tensor-core) leg")
    ap.add_argument("--dense-only", action="store_true", help="profiling runs: only the cfg3 deep-unit leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the cfg4 strong-scaling leg")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the cfg5 sparsity sweep")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the cfg1 CLI leg (stock vs patched l-giremi)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
