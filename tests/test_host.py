"""CPU-side tests: the C-ABI library loads and exports what include/lgmi.h
declares, fails loudly without a device, and the host logic (encoder,
partitioner, synthetic generator) agrees with the oracle.  No compute call is
made here -- those live in test_gpu_parity.py."""
import ctypes
import os
import re
import subprocess
import tempfile
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_mismatches, has_gpu

sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import c_oracle  # noqa: E402
from fuzz import random_mismatches  # noqa: E402


# --------------------------------------------------------------------------- ABI
def header_functions():
    src = open(os.path.join(ROOT, "include", "lgmi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgmi_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(liblgmi_path, lg):
    declared = header_functions()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", liblgmi_path], capture_output=True, text=True, check=True)
    exported = set(line.split()[-1] for line in out.stdout.splitlines() if " T " in line)
    missing = [f for f in declared if f not in exported]
    assert not missing, missing
    # the ctypes binding declares the same set
    binding = importlib_lib().EXPORTS
    assert sorted(binding) == declared
    # nothing but the ABI leaks out of the library
    stray = [s for s in exported if not s.startswith("lgmi_")]
    assert not stray, stray


def importlib_lib():
    import importlib
    return importlib.import_module("l-giremi_b200._lib")


def test_abi_struct_layouts(lg):
    L = importlib_lib()
    assert L.UNIT_DESC.itemsize == 24 and L.UNIT_DESC.fields["site_off"][1] == 20
    assert L.PAIR_REC.itemsize == 16 and L.PAIR_REC.fields["mi"][1] == 8
    assert ctypes.sizeof(L.Result) == 136
    # ... and the C compiler agrees with the ctypes mirror, field by field
    fields = [f[0] for f in L.Result._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "lgmi.h"\nint main(void){printf("%zu", sizeof(lgmi_result));' + \
        "".join('printf(" %%zu", offsetof(lgmi_result, %s));' % f for f in fields) + "return 0;}"
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "abi.c"), os.path.join(tmp, "abi")
        with open(src, "w") as fh:
            fh.write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(L.Result)] + [getattr(L.Result, f).offset for f in fields]


def test_version_and_pure_host_entry_points(lg):
    lib = importlib_lib().load()
    assert lib.lgmi_version() == 100
    assert lib.lgmi_unit_cost(50, 200) == 1225 * 4
    assert lib.lgmi_unit_cost(1, 200) == 0
    assert lib.lgmi_unit_cost(2000, 100000) == 1999000 * 1563


def test_pipeline_halves_reject_a_null_handle(lg):
    """lgmi_pipeline_begin* / lgmi_pipeline_finish without a pipeline: LGMI_ERR_ARG, no device touched."""
    import ctypes as C
    lib = importlib_lib().load()
    res = importlib_lib().Result()
    assert lib.lgmi_pipeline_begin(None, None, None, 6, 0) == -2
    assert lib.lgmi_pipeline_begin_packed(None, None, None, 6, 0) == -2
    assert lib.lgmi_pipeline_collect(None) == -2
    assert lib.lgmi_pipeline_finish(None, C.byref(res)) == -2
    assert lib.lgmi_pipeline_step_packed(None, None, None, 6, 0, C.byref(res)) == -2


@pytest.mark.skipif(has_gpu(), reason="checks the no-device error path")
def test_no_device_is_a_loud_error(lg):
    """No CPU fallback: creating a context without a GPU raises."""
    with pytest.raises(lg.LgmiError) as e:
        lg.Context(0)
    assert e.value.code == -5
    assert "no CUDA device" in str(e.value) or "sm_" in str(e.value)
    site = {'ref': 'A', 'type': 'het_snp', 'depth': {'A': 6, 'C': 6},
            'nt': {'A': ['r%d' % k for k in range(6)], 'C': ['r%d' % k for k in range(6, 12)]}}
    with pytest.raises(lg.LgmiError):
        lg.mismatch_pair_mutual_info({10: site, 20: site}, 6)
    with pytest.raises(lg.LgmiError):
        lg.ecdf([0.1, 0.2])(0.15)


def test_product_never_imports_the_oracle():
    """The shipped package must not import, load or execute oracle/ or any CPU
    MI implementation (scikit-learn): it fails without its CUDA library."""
    pkg = os.path.join(ROOT, "l-giremi_b200")
    bad = re.compile(r"^\s*(from|import)\s+(oracle|c_oracle|ref_port|sklearn|scipy)\b|liboracle|oracle_mi|"
                     r"sys\.path.*oracle", re.M)
    n = 0
    for dirpath, _dirs, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert not bad.search(text), os.path.join(dirpath, name)
                n += 1
    assert n >= 8


# --------------------------------------------------------------------------- encoder
def test_encoder_labels_equal_oracle_labels(lg, golden):
    rng = np.random.default_rng(11)
    cases = [golden_mismatches(u) for u in golden("units_fuzz.json")]
    cases += [random_mismatches(rng) for _ in range(150)]
    for m in cases:
        eu = lg.encode_mismatches(m)
        coded = c_oracle.code_mismatches(m)
        want, bad = c_oracle.labels_of(coded)
        got = eu.labels.astype(np.int16)
        got[got == 255] = -1
        # the two may number the reads differently only if names were seen in a different order
        assert got.shape == want.shape
        assert np.array_equal(got, want.astype(np.int16))
        assert sorted(eu.bad_sites) == np.nonzero(bad)[0].tolist()
        assert eu.positions == coded["positions"] and eu.types == coded["types"]


def test_encoder_does_not_mutate_input(lg):
    m = random_mismatches(np.random.default_rng(5))
    import copy
    before = copy.deepcopy(m)
    lg.encode_mismatches(m)
    assert m == before


def test_plane_packing_roundtrip(lg):
    enc = __import__("importlib").import_module("l-giremi_b200.encode")
    rng = np.random.default_rng(7)
    for R in (1, 31, 32, 33, 127, 128, 129, 200, 1000):
        S = int(rng.integers(1, 7))
        labels = rng.choice(np.array([0, 1, 2, 255], dtype=np.uint8), size=(S, R), p=[0.05, 0.3, 0.4, 0.25])
        planes = enc.pack_labels(labels)
        W = enc.row_words(R)
        assert planes.shape == (S, 3, W) and W % 4 == 0 and W * 32 >= R
        bits = np.unpackbits(planes.view(np.uint8).reshape(S, 3, W * 4), axis=-1, bitorder="little")
        assert np.array_equal(bits[:, 0, :R], labels == 2)
        assert np.array_equal(bits[:, 1, :R], labels == 1)
        assert np.array_equal(bits[:, 2, :R], labels != 255)
        assert not bits[:, :, R:].any()                      # pad bits are zero


def test_pack_units_descriptor_invariants(lg):
    rng = np.random.default_rng(8)
    ms = [random_mismatches(rng) for _ in range(9)] + [{}]
    pb = lg.encode_batch(ms)
    assert pb.n_units == 10
    off = 0
    soff = 0
    for u in pb.units:
        assert u["plane_off"] == off and u["plane_off"] % 4 == 0 and u["site_off"] == soff
        assert u["row_words"] == 4 * ((u["n_reads"] + 127) // 128)
        off += 3 * int(u["n_sites"]) * int(u["row_words"])
        soff += int(u["n_sites"])
    assert off == pb.planes.size and soff == pb.n_sites
    sub = pb.subset([3, 1])
    assert sub.n_units == 2 and sub.units["plane_off"][0] == 0
    assert sub.site_types(0) == pb.site_types(3)


# --------------------------------------------------------------------------- synthetic generator
def test_synth_dict_and_planes_describe_the_same_unit(lg, golden):
    synth = __import__("importlib").import_module("l-giremi_b200.synth")
    sb = synth.make_uniform(11, 3, 12, 48, 0.5)
    pb = sb.plane_batch()
    for g in range(3):
        eu = lg.encode_mismatches(sb.mismatches(g))
        direct = sb.encoded(g)
        # dict route numbers reads by first appearance; compare as multisets of read columns
        a = sorted(map(bytes, eu.labels.T))
        b = sorted(map(bytes, direct.labels.T[(direct.labels != 255).any(axis=0)]))
        assert a == b
        assert eu.types == direct.types and eu.positions == direct.positions
    # the committed golden units were drawn by this generator: it must be reproducible
    unit = golden("units_synth.json")[0]
    assert golden_mismatches(unit) == {p: {**s} for p, s in
                                       ((p, dict(ref=s['ref'], type=s['type'], depth=s['depth'], nt=s['nt'],
                                                 neighbor=[], up='A', down='C'))
                                        for p, s in sb.mismatches(0).items())}
    same = synth.make_uniform_planes(11, 3, 12, 48, 0.5)
    assert np.array_equal(same.planes, pb.planes) and np.array_equal(same.site_flags, pb.site_flags)


def test_synth_shapes_for_cfg2_and_heavy_tail(lg):
    synth = __import__("importlib").import_module("l-giremi_b200.synth")
    pb = synth.make_uniform_planes(20261020, 40, 50, 200, 0.5)
    assert pb.n_candidates == 40 * 1225 and pb.planes.size == 40 * 3 * 50 * 8
    flags = pb.site_flags & 3
    assert 0.03 < (flags == 2).mean() < 0.2
    hp, _ = synth.make_heavy_tail(20261022, 60)
    assert hp.units["n_sites"].min() >= 2 and hp.units["n_sites"].max() <= 1000
    assert hp.units["n_reads"].min() >= 6 and hp.units["n_reads"].max() <= 20000


# --------------------------------------------------------------------------- partitioner
def test_lpt_partition_properties(lg):
    rng = np.random.default_rng(9)
    cost = rng.lognormal(5, 1.5, 5000).astype(np.uint64)
    for nb in (1, 2, 4, 8):
        bin_of, load = lg.partition_lpt(cost, nb)
        assert bin_of.min() >= 0 and bin_of.max() < nb
        assert np.array_equal(np.bincount(bin_of, weights=cost.astype(np.float64), minlength=nb).astype(np.uint64), load)
        # LPT guarantee: max load <= mean + largest item
        assert load.max() <= cost.sum() / nb + cost.max()
        assert load.max() - load.min() <= cost.max()
        again, _ = lg.partition_lpt(cost, nb)
        assert np.array_equal(again, bin_of)                 # deterministic


def test_lpt_matches_a_plain_python_greedy(lg):
    rng = np.random.default_rng(10)
    cost = rng.integers(0, 1000, 300).astype(np.uint64)
    nb = 4
    order = sorted(range(len(cost)), key=lambda k: (-int(cost[k]), k))
    load = [0] * nb
    want = [0] * len(cost)
    for k in order:
        b = min(range(nb), key=lambda x: (load[x], x))
        want[k] = b
        load[b] += int(cost[k])
    bin_of, got_load = lg.partition_lpt(cost, nb)
    assert bin_of.tolist() == want and got_load.tolist() == load


def test_unit_costs_formula(lg):
    L = importlib_lib()
    units = np.zeros(4, dtype=L.UNIT_DESC)
    units["n_sites"] = [0, 1, 50, 2000]
    units["n_reads"] = [10, 10, 200, 100000]
    lib = L.load()
    want = [lib.lgmi_unit_cost(int(s), int(r)) for s, r in zip(units["n_sites"], units["n_reads"])]
    assert lg.unit_costs(units).tolist() == want == [0, 0, 4900, 1999000 * 1563]


def test_product_sources_carry_no_compiled_out_experiments():
    """Round-1's #ifdef experiments (means one unit behind, per-barrier cycle counters) live under
    tools/experiments/ as text; the product headers have no dead branches."""
    import glob
    for path in glob.glob(os.path.join(ROOT, "l-giremi_b200", "csrc", "*")):
        if path.endswith((".cu", ".cuh", ".inl")):
            text = open(path).read()
            assert "LGMI_HET_DEFERRED_MEANS" not in text and "LGMI_PHASE_CLOCKS" not in text, path



def test_tight_two_plane_form_is_the_padded_form_without_the_padding(lg):
    """PlaneBatch.packed2(tight=True) (LGMI_MODE_TIGHT_INPUT): rows of ceil(R/32) words, units back to back;
    word for word the padded two-plane form minus its all-zero tail words."""
    import importlib
    synth = importlib.import_module("l-giremi_b200.synth")
    pb, _ = synth.make_heavy_tail(3, 60, s_max=80, r_max=700)
    padded, tight = pb.packed2(), pb.packed2(tight=True)
    S = pb.units['n_sites'].astype(int)
    W = pb.units['row_words'].astype(int)
    R = pb.units['n_reads'].astype(int)
    assert padded.size == pb.planes.size // 3 * 2
    assert tight.size == int((2 * S * ((R + 31) // 32)).sum()) < padded.size
    off = 0
    for k in range(pb.n_units):
        wt, w = (R[k] + 31) // 32, W[k]
        a = padded[int(pb.units['plane_off'][k]) // 3 * 2:][:2 * S[k] * w].reshape(S[k], 2, w)
        b = tight[off:off + 2 * S[k] * wt].reshape(S[k], 2, wt)
        off += 2 * S[k] * wt
        assert np.array_equal(a[:, :, :wt], b) and not a[:, :, wt:].any()
        # 2 bits per read: 01 major, 10 minor, 11 other
        M, m, C = (pb.planes[int(pb.units['plane_off'][k]):][:3 * S[k] * w].reshape(S[k], 3, w)[:, q] for q in range(3))
        assert np.array_equal(a[:, 0] & ~a[:, 1], M) and np.array_equal(a[:, 1] & ~a[:, 0], m)
        assert np.array_equal(a[:, 0] | a[:, 1], C)
    assert off == tight.size


class _FakePipeline:
    """Records the calls stream_steps makes and enforces the library's state rules (one step in flight per
    pipeline: a second begin, or a collect / finish without a begin, is an error)."""

    def __init__(self, log, name):
        self.log, self.name, self.step, self.collected = log, name, None, False

    def begin(self, min_common, mode, planes, site_flags, packed=False, tight=False):
        assert self.step is None, "begin on a pipeline whose step has not been finished"
        self.step, self.collected = int(planes[0]), False
        self.log.append(('begin', self.step, self.name))

    def collect(self):
        assert self.step is not None and not self.collected
        self.collected = True
        self.log.append(('collect', self.step, self.name))

    def finish(self, copy=True):
        assert self.step is not None
        k, self.step = self.step, None
        self.log.append(('finish', k, self.name))
        return k


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 6])
def test_stream_steps_keeps_depth_steps_in_flight(lg, depth):
    """stream_steps / stream_schedule: every step begun, (collected,) finished once and yielded in input order,
    never more than `depth` in flight, the downloads of the next step queued before the previous one is waited
    for (depth >= 3), and the calls are exactly stream_schedule's."""
    for n in (0, 1, 2, 3, 5, 11):
        log = []
        pipes = [_FakePipeline(log, q) for q in range(depth)]
        inputs = ((np.array([k], np.uint32), np.zeros(1, np.uint8)) for k in range(n))
        out = list(lg.stream_steps(pipes, inputs, 6))
        assert out == [(k, k) for k in range(n)]
        assert [(op, k) for op, k, _q in log] == list(lg.stream_schedule(n, depth))
        assert all(q == k % depth for _op, k, q in log)
        in_flight = 0
        for op, k, _q in log:
            in_flight += (op == 'begin') - (op == 'finish')
            assert 0 <= in_flight <= depth
        for op in ('begin', 'finish'):
            assert [k for o, k, _q in log if o == op] == list(range(n))
        if depth >= 3:
            pos = {(op, k): t for t, (op, k, _q) in enumerate(log)}
            for k in range(n - 1):
                if ('collect', k + 1) in pos:                # queued before the host waits for step k
                    assert pos[('collect', k + 1)] < pos[('finish', k)]
            assert sum(1 for o, _k, _q in log if o == 'collect') == max(0, n - (depth - 2)) if n else True
