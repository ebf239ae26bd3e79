"""CPU-only: the kernels of morefusion_b200/csrc/mapping.cu executed serially on the host
(tests/emu/mapping_emu.cpp, g++ -DMF_HOST_EMU) against oracle/octomap.py, and the host logic of
``MultiInstanceOctreeMapping`` (table growth, overflow errors, API contract) driven through the
same emulation.  The GPU run of the same kernels is tests/test_mapping_gpu.py."""

import contextlib
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import octomap as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

c_f, c_i, c_i64, c_p, c_d, c_u32 = (ctypes.c_float, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                     ctypes.c_double, ctypes.c_uint32)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "mapping_emu.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17",
                    "-I" + os.path.join(ROOT, "tests", "emu"),
                    os.path.join(ROOT, "tests", "emu", "mapping_emu.cpp"), "-o", so], check=True)
    L = ctypes.CDLL(so)
    from morefusion_b200 import _lib
    names = ("integrate", "integrate_labelled", "update_points", "query_grids", "rehash")
    for name in names:
        res, args = _lib.SIGNATURES["mf_map_" + name]
        fn = getattr(L, "emu_map_" + name)
        fn.restype = res
        fn.argtypes = args[:-1]                                  # no stream

    class Adapter:
        pass
    a = Adapter()
    for name in names:
        setattr(a, "mf_map_" + name, (lambda f: (lambda *args: f(*args[:-1])))(getattr(L, "emu_map_" + name)))
    return a


def make_scene(seed, H=20, W=28):
    """A small depth frame: a box-ish foreground object in front of a tilted background plane."""
    rs = np.random.RandomState(seed)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    f = 40.0
    z = 0.55 + 0.1 * (u / W) + 0.01 * rs.rand(H, W)
    fg = (np.abs(u - W / 2) < W / 5) & (np.abs(v - H / 2) < H / 4)
    z = np.where(fg, 0.38 + 0.004 * rs.rand(H, W), z)
    pcd = np.stack([(u - W / 2) * z / f, (v - H / 2) * z / f, z], -1).astype(np.float32)
    pcd[rs.rand(H, W) < 0.05] = np.nan
    return pcd, fg


@pytest.fixture()
def emu_mapping(emu, monkeypatch):
    from morefusion_b200 import _lib
    from morefusion_b200.contrib.multi_instance_octree_mapping import MultiInstanceOctreeMapping
    monkeypatch.setattr(_lib, "lib", lambda: emu)
    monkeypatch.setattr(_lib, "stream", lambda: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)

    class EmuMapping(MultiInstanceOctreeMapping):           # host logic unchanged; no CUDA objects
        def _dev_ctx(self):
            return contextlib.nullcontext()

        def _read_back(self):
            self._host_counters = self._counters.clone()

            class Done:
                def synchronize(self):
                    pass

                def query(self):
                    return True
            self._pending = Done()

    return EmuMapping


def build_pair(cls, scans, capacity=1 << 12, device="cpu"):
    ours = cls(device=device, capacity=capacity)
    ref = oc.MultiInstanceOctreeMapping()
    for ins, pitch in ((3, 0.008), (7, 0.011), (0, 0.02)):
        ours.initialize(ins, pitch=pitch)
        ref.initialize(ins, pitch=pitch)
    for ins, mask, pcd, origin in scans:
        ours.integrate(ins, mask, pcd, origin)
        ref.integrate(ins, mask, pcd, origin)
    return ours, ref


def scans_for(seed):
    pcd, fg = make_scene(seed)
    pcd2, fg2 = make_scene(seed + 100)
    org2 = (0.013, -0.021, 0.004)
    left = np.zeros_like(fg)
    left[:, : fg.shape[1] // 2] = True
    return [(3, fg, pcd, (0, 0, 0)), (0, ~fg, pcd, (0, 0, 0)), (7, fg & left, pcd, (0, 0, 0)),
            (3, fg2, pcd2 + np.float32(org2), org2), (0, ~fg2, pcd2 + np.float32(org2), org2)]


def assert_cells_equal(ours, ref):
    for ins in ref.instance_ids:
        a, b = ours.cells(ins), ref._octrees[ins].cells
        assert set(a) == set(b), (ins, len(a), len(b))
        for k in b:
            assert a[k] == b[k], (ins, k, a[k], b[k])            # float32 log-odds bit for bit


@pytest.mark.parametrize("seed", [0, 1])
def test_scan_integration_bit_exact(emu_mapping, seed):
    ours, ref = build_pair(emu_mapping, scans_for(seed))
    assert_cells_equal(ours, ref)
    n = sum(len(t.cells) for t in ref._octrees.values())
    assert ours.n_cells() == n and n > 1500


def test_target_grids_match_oracle(emu_mapping):
    ours, ref = build_pair(emu_mapping, scans_for(0))
    pcd, fg = make_scene(0)
    for target, pitch in ((3, 0.008), (7, 0.011)):
        center = np.nanmedian(pcd[fg], axis=0)
        origin = center - (16 / 2 - 0.5) * pitch
        got = ours.get_target_grids(target, dimensions=(16, 16, 16), pitch=pitch, origin=origin)
        want = ref.get_target_grids(target, dimensions=(16, 16, 16), pitch=pitch, origin=origin)
        for g, w in zip(got, want):
            assert g.dtype == np.float32 and g.shape == (16, 16, 16)
            np.testing.assert_array_equal(g > 0, w > 0)
            np.testing.assert_allclose(g, w, rtol=0, atol=1e-7)
        assert (want[0] > 0).sum() > 20 and (want[2] > 0).sum() > 100
    # several targets, non-cubic grid, in one launch
    gt, gn, ge = ours.get_target_grids_batch([7, 3], dimensions=(6, 9, 5), pitches=[0.011, 0.02],
                                             origins=[[-0.02, -0.03, 0.36], [-0.05, -0.05, 0.3]])
    for t, (tid, pitch, org) in enumerate(((7, 0.011, [-0.02, -0.03, 0.36]), (3, 0.02, [-0.05, -0.05, 0.3]))):
        want = ref.get_target_grids(tid, dimensions=(6, 9, 5), pitch=pitch, origin=org)
        for g, w in zip((gt[t], gn[t], ge[t]), want):
            np.testing.assert_allclose(g.numpy(), w, rtol=0, atol=1e-7)


def test_update_points_and_pcds(emu_mapping):
    ours, ref = build_pair(emu_mapping, scans_for(1)[:2])
    rs = np.random.RandomState(5)
    occupied = np.concatenate([rs.uniform(-0.05, 0.05, (300, 3)) + [0, 0, 0.4],
                               np.repeat([[0.001, 0.002, 0.4]], 9, 0),       # 9 hits on one cell: clamps
                               [[np.nan, 0, 0], [1e9, 0, 0]]])
    ours.update(3, occupied)
    ref.update(3, occupied[:-2])                  # the oracle's int() rejects NaN / out of range rows
    assert_cells_equal(ours, ref)
    for kw in ({}, dict(aabb_min=(-0.03, -0.03, 0.3), aabb_max=(0.03, 0.04, 0.45))):
        a = ours.get_target_pcds(3, **kw)
        b = ref.get_target_pcds(3, **kw)
        for x, y in zip(a, b):
            assert x.dtype == np.float64
            np.testing.assert_array_equal(x, y)


def test_table_growth_keeps_every_cell(emu_mapping):
    ours, ref = build_pair(emu_mapping, scans_for(0), capacity=1 << 12)
    assert ours._cap > (1 << 12)                               # grew (x4 when a quarter full)
    assert_cells_equal(ours, ref)


def test_overflow_is_reported_not_silent(emu_mapping):
    pcd, fg = make_scene(0)
    m = emu_mapping(device="cpu", capacity=64)
    m.initialize(1, pitch=0.008)
    m.integrate(1, fg, pcd)                                     # far more than 64 cells
    with pytest.raises(RuntimeError, match="overflowed"):
        m.integrate(1, fg, pcd)
    m = emu_mapping(device="cpu", capacity=64)
    m.initialize(1, pitch=0.008)
    m.integrate(1, fg, pcd)
    with pytest.raises(RuntimeError, match="overflowed"):       # queries check too
        m.get_target_grids(1, dimensions=(4, 4, 4), pitch=0.01, origin=(0, 0, 0))


def test_api_contract(emu_mapping):
    m = emu_mapping(device="cpu", capacity=1 << 10)
    m.initialize(5, pitch=0.01)
    assert m.instance_ids == [5]
    with pytest.raises(ValueError, match="already exists"):      # multi_instance_octree_mapping.py:17-18
        m.initialize(5, pitch=0.01)
    with pytest.raises(KeyError):
        m.integrate(6, np.zeros((2, 2), bool), np.zeros((2, 2, 3), np.float32))
    with pytest.raises(AssertionError):                          # :55-58
        m.get_target_grids(5, dimensions=(4, 4, 4), pitch=0.01, origin=(np.nan, 0, 0))
    with pytest.raises(AssertionError):
        m.get_target_grids(5, dimensions=(4, 4, 4), pitch=-1.0, origin=(0, 0, 0))
    g = m.get_target_grids(5, dimensions=(4, 4, 4), pitch=0.01, origin=(0, 0, 0))
    assert all((x == 0).all() for x in g)                        # nothing integrated: all unknown


def run_golden(cls, g, **kw):
    m = cls(**kw)
    for ins, pitch in zip(g["instance_ids"], g["pitches"]):
        m.initialize(int(ins), pitch=float(pitch))
    for n in range(2):
        for ins in g["instance_ids"]:
            m.integrate(int(ins), g[f"label{n}"] == ins, g[f"pcd{n}"], origin=g[f"origin{n}"])
    return m


def check_golden(m, g, to_np=lambda a: a):
    for ins in g["instance_ids"]:
        cells = m.cells(int(ins))
        keys = g[f"cells_keys_{ins}"]
        assert len(cells) == len(keys), (ins, len(cells), len(keys))
        got = np.array([cells[tuple(k)] for k in keys], np.float32)
        assert np.array_equal(got, g[f"cells_logodds_{ins}"]), ins        # float32 log-odds bit for bit
    for tid, pitch in ((1, 0.006), (3, 0.005)):
        grids = m.get_target_grids(tid, dimensions=(32, 32, 32), pitch=pitch, origin=g[f"grid_origin_{tid}"])
        for name, a in zip(("target", "nontarget", "empty"), grids):
            w = g[f"ref_grid_{name}_{tid}"]
            assert a.dtype == np.float32
            np.testing.assert_array_equal(a > 0, w > 0)                   # occupancy pattern exact
            np.testing.assert_allclose(a, w, rtol=0, atol=1e-7)          # exp(): 1 ulp of float32
    occ, emp = m.get_target_pcds(2)
    np.testing.assert_array_equal(occ, g["ref_pcd_occupied_2"])
    np.testing.assert_array_equal(emp, g["ref_pcd_empty_2"])


def run_golden_update(make, g):
    m = make()
    m.initialize(1, pitch=0.006)
    m.initialize(0, pitch=0.012)
    m.integrate(1, g["label"] == 1, g["pcd"], origin=g["origin"])
    m.integrate(0, g["label"] != 1, g["pcd"], origin=g["origin"])
    m.update(1, g["occupied"])
    return m


def test_golden_reference_update(emu_mapping):
    """``update`` (updateNodes row by row, clamping after 12 hits on one cell) as the reference's
    own class performs it over the restated OcTree (tests/golden/octree_mapping_update.npz)."""
    from conftest import golden
    g = golden("octree_mapping_update")
    for m in (run_golden_update(lambda: emu_mapping(device="cpu", capacity=1 << 14), g),
              run_golden_update(oc.MultiInstanceOctreeMapping, g)):
        cells = m.cells(1) if hasattr(m, "cells") else m._octrees[1].cells
        keys = g["cells_keys_1"]
        assert len(cells) == len(keys)
        assert np.array_equal(np.array([cells[tuple(k)] for k in keys], np.float32), g["cells_logodds_1"])
        grids = m.get_target_grids(1, dimensions=(16, 16, 16), pitch=0.006, origin=g["grid_origin"])
        for name, a in zip(("target", "nontarget", "empty"), grids):
            np.testing.assert_array_equal(a > 0, g[f"ref_grid_{name}"] > 0)
            np.testing.assert_allclose(a, g[f"ref_grid_{name}"], rtol=0, atol=1e-7)
    assert g["cells_logodds_1"].max() == np.float32(oc.logodds(0.971))      # the clamp was reached


def test_golden_reference_run(emu_mapping):
    from conftest import golden
    g = golden("octree_mapping")
    check_golden(run_golden(emu_mapping, g, device="cpu", capacity=1 << 14), g)


def test_labelled_frame_equals_per_instance_scans(emu_mapping):
    from conftest import golden
    g = golden("octree_mapping")
    m = emu_mapping(device="cpu", capacity=1 << 18)
    for ins, pitch in zip(g["instance_ids"], g["pitches"]):
        m.initialize(int(ins), pitch=float(pitch))
    for n in range(2):
        lab = g[f"label{n}"].copy()
        lab[0, :5] = 77                                           # a label nobody initialised: skipped
        lab[1, :5] = -3
        pcd = g[f"pcd{n}"].copy()
        pcd[0, :5] = np.nan
        pcd[1, :5] = np.nan
        m.integrate_labels(lab, pcd, origin=g[f"origin{n}"])
    ref = run_golden(emu_mapping, dict(g, pcd0=_nan_rows(g["pcd0"]), pcd1=_nan_rows(g["pcd1"])),
                     device="cpu", capacity=1 << 18)
    for ins in g["instance_ids"]:
        assert m.cells(int(ins)) == ref.cells(int(ins))
    assert m._scan == 2 and ref._scan == 8                        # 2 launches pairs instead of 8


def test_labelled_frame_with_negative_and_sparse_instance_ids(emu_mapping):
    """Instance ids as the ROS node numbers them (-1 = background, arbitrary positive ids): the
    label look-up table covers [min id, max id] and skips everything else."""
    pcd, fg = make_scene(2)
    label = np.where(fg, 41, -1).astype(np.int64)
    label[:2] = -2                                            # "uncertain" pixels: never integrated
    label[2, :4] = 17                                         # inside the table's range, not an instance
    a, b = emu_mapping(device="cpu", capacity=1 << 14), emu_mapping(device="cpu", capacity=1 << 14)
    for m in (a, b):
        m.initialize(41, pitch=0.008)
        m.initialize(-1, pitch=0.02)
    a.integrate_labels(label, pcd)
    b.integrate(41, label == 41, pcd)
    b.integrate(-1, label == -1, pcd)
    for ins in (41, -1):
        assert a.cells(ins) == b.cells(ins) and len(a.cells(ins)) > 100
    with pytest.raises(ValueError, match="integer"):
        a.integrate_labels(label.astype(np.float32), pcd)


def _nan_rows(pcd):
    pcd = pcd.copy()
    pcd[0, :5] = np.nan
    pcd[1, :5] = np.nan
    return pcd


@pytest.mark.parametrize("seed", range(6))
def test_random_scans_negative_coordinates_and_odd_pitches(emu_mapping, seed):
    """Random point sets on both sides of every axis, sensor origins off the lattice, pitches that
    are not exactly representable: the cell keys, the DDA and the once-per-scan rule against the
    oracle, bit for bit, after three scans per instance."""
    rs = np.random.RandomState(100 + seed)
    ours, ref = emu_mapping(device="cpu", capacity=1 << 15), oc.MultiInstanceOctreeMapping()
    pitches = {4: float(rs.uniform(0.003, 0.02)), 9: 0.0075, 0: float(rs.uniform(0.01, 0.03))}
    for ins, pitch in pitches.items():
        ours.initialize(ins, pitch=pitch)
        ref.initialize(ins, pitch=pitch)
    for scan in range(3):
        origin = rs.uniform(-0.05, 0.05, 3)
        pcd = rs.uniform(-0.25, 0.25, (10, 12, 3)).astype(np.float32)
        pcd[rs.rand(10, 12) < 0.1] = np.nan
        if scan == 1:
            pcd[0, 0] = origin.astype(np.float32)                # zero-length ray
            pcd[0, 1] = [1e6, 0, 0]                              # end point outside the key range
        label = rs.choice([4, 9, 0, 5], size=(10, 12))           # 5 is never initialised
        for ins in pitches:
            ours.integrate(ins, label == ins, pcd, origin)
            ref.integrate(ins, label == ins, pcd, origin)
    assert_cells_equal(ours, ref)
    for tid in (4, 9):
        org = rs.uniform(-0.2, -0.1, 3)
        got = ours.get_target_grids(tid, dimensions=(12, 12, 12), pitch=0.02, origin=org)
        want = ref.get_target_grids(tid, dimensions=(12, 12, 12), pitch=0.02, origin=org)
        for g, w in zip(got, want):
            np.testing.assert_array_equal(g > 0, w > 0)
            np.testing.assert_allclose(g, w, rtol=0, atol=1e-7)


def test_ray_walk_against_oracle_dense(emu_mapping):
    """The DDA over many random rays, including axis-aligned, zero-length and far rays, from an
    origin that sits exactly on cell borders."""
    rs = np.random.RandomState(3)
    ends = rs.uniform(-0.6, 0.6, (400, 3)).astype(np.float32)
    ends[:20, 0] = 0.25                                          # shared coordinates with the origin
    ends[20:40, 1:] = np.float32([-0.125, 0.0625])
    ends[40] = [0.25, -0.125, 0.0625]                            # same cell as the origin
    origin = np.float32([0.25, -0.125, 0.0625])
    res = 0.0125
    m = emu_mapping(device="cpu", capacity=1 << 17)
    m.initialize(2, pitch=res)
    m.integrate(2, np.ones((20, 20), bool), ends.reshape(20, 20, 3), origin=origin)
    t = oc.OcTree(res)
    t.insertPointCloud(ends, origin=origin)
    got = m.cells(2)
    assert set(got) == set(t.cells)
    assert all(got[k] == t.cells[k] for k in got)
    walked = set()
    for e in ends:
        walked.update(t.computeRayKeys(origin, e))
    assert len(walked) > 5000 and walked <= set(got)
