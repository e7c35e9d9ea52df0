"""Multi-GPU paths on real devices (the host logic alone is tests/test_shard_gloo.py).

  * multigpu.DevicePool -- ONE process owning several devices, what the patched `l-giremi` parent
    uses -- against a single submit, bit for bit.  With one GPU in the box the pool is built over
    two contexts on device 0, which exercises the same partition / thread / merge path.
  * one process per GPU under torch.distributed.run with the nccl gather (tools/shard_check.py),
    skipped when the box has fewer than two GPUs."""
import importlib
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

synth = importlib.import_module("l-giremi_b200.synth")


def test_device_pool_equals_single_submit(lg, gpu_ctx):
    n = lg.device_count()
    assert n >= 1
    devices = list(range(n)) if n >= 2 else [0, 0]
    pool = lg.DevicePool(devices)
    try:
        pb, _ = synth.make_heavy_tail(20261034, 240, s_max=200, r_max=2500)
        for mode in (lg.MODE_HET_ONLY | lg.MODE_SKIP_NONHET, lg.MODE_ALL_PAIRS | lg.MODE_EMIT_COUNTS):
            want = lg.mi_step_batched(pb, 6, mode, ctx=gpu_ctx)
            got = pool.run(pb, 6, mode)
            assert isinstance(got, lg.MergedResult) and len(got.shard_units) == len(devices)
            assert sum(got.shard_units) == pb.n_units
            assert np.array_equal(got.records, want.records)
            assert np.array_equal(got.site_mean, want.site_mean, equal_nan=True)
            assert np.array_equal(got.site_cnt, want.site_cnt)
            assert np.array_equal(got.unit_rec_off, want.unit_rec_off)
            if mode & lg.MODE_EMIT_COUNTS:
                assert np.array_equal(got.counts, want.counts)
        # a batch smaller than two units per device is one submit on the first device
        tiny, _ = synth.make_heavy_tail(5, 3, s_max=40, r_max=200)
        assert isinstance(pool.run(tiny, 6, lg.MODE_HET_ONLY), lg.StepResult)
    finally:
        pool.close()


def test_default_device_follows_the_environment(lg, monkeypatch):
    monkeypatch.setenv("LGMI_DEVICE", "3")
    assert lg.default_device() == 3
    monkeypatch.delenv("LGMI_DEVICE")
    monkeypatch.setenv("LOCAL_RANK", "2")
    assert lg.default_device() == 2
    monkeypatch.delenv("LOCAL_RANK")
    assert lg.default_device() == 0                      # not a pool worker: the first GPU


@pytest.mark.timeout(600)
def test_one_process_per_gpu_gathered_result_equals_single_gpu(lg):
    n = lg.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs (run with gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "shard_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=550)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
    assert line["ok"] and line["n_gpus"] == world
