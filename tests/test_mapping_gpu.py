"""GPU parity of the occupancy-grid producer: csrc/mapping.cu through the C ABI
against oracle/octomap.py, the committed fixture produced by the reference's own class
(tests/golden/octree_mapping.npz) and, at the full 640x480 frame size, size-independent properties
(independence of thread order, scan idempotence of the stamps, batch query == single queries)."""

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import octomap as oc
from test_mapping_emu import (assert_cells_equal, build_pair, check_golden, make_scene, run_golden,
                              scans_for)

pytestmark = pytest.mark.gpu


@pytest.fixture()
def Mapping(cuda_device):
    from morefusion_b200.contrib import MultiInstanceOctreeMapping
    return MultiInstanceOctreeMapping


def test_golden_reference_run_gpu(Mapping):
    g = golden("octree_mapping")
    check_golden(run_golden(Mapping, g, capacity=1 << 14), g)          # 1<<14: also exercises growth


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_scans_and_grids_vs_oracle(Mapping, seed):
    ours, ref = build_pair(Mapping, scans_for(seed), capacity=1 << 16, device="cuda")
    assert_cells_equal(ours, ref)
    pcd, fg = make_scene(seed)
    center = np.nanmedian(pcd[fg], axis=0)
    for target, pitch in ((3, 0.008), (7, 0.011)):
        origin = center - 7.5 * pitch
        got = ours.get_target_grids(target, dimensions=(16, 16, 16), pitch=pitch, origin=origin)
        want = ref.get_target_grids(target, dimensions=(16, 16, 16), pitch=pitch, origin=origin)
        for a, w in zip(got, want):
            np.testing.assert_array_equal(a > 0, w > 0)
            np.testing.assert_allclose(a, w, rtol=0, atol=1e-7)


def test_update_points_vs_oracle(Mapping):
    ours, ref = build_pair(Mapping, scans_for(1)[:2], capacity=1 << 16, device="cuda")
    rs = np.random.RandomState(5)
    occupied = np.concatenate([rs.uniform(-0.05, 0.05, (3000, 3)) + [0, 0, 0.4],
                               np.repeat([[0.001, 0.002, 0.4]], 9, 0), [[np.nan, 0, 0], [1e9, 0, 0]]])
    ours.update(3, occupied)
    ref.update(3, occupied[:-2])
    assert_cells_equal(ours, ref)
    for x, y in zip(ours.get_target_pcds(3), ref.get_target_pcds(3)):
        np.testing.assert_array_equal(x, y)


def full_frame(seed, H=480, W=640):
    """640x480 synthetic RGB-D frame: 8 boxes on a tilted table (cfg2/cfg5 geometry)."""
    rs = np.random.RandomState(seed)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    fx = 600.0
    z = (0.9 + 0.2 * (v / H)).astype(np.float64)
    label = np.zeros((H, W), np.int32)
    for ins in range(1, 9):
        cu, cv = 80 + 60 * ins, 140 + 25 * (ins % 4)
        m = (np.abs(u - cu) < 28) & (np.abs(v - cv) < 36)
        z = np.where(m, 0.55 + 0.03 * (ins % 3) + 0.01 * ((u - cu) / 28.0) ** 2, z)
        label[m] = ins
    z = z + 0.003 * rs.randn(H, W)
    pcd = np.stack([(u - W / 2) * z / fx, (v - H / 2) * z / fx, z], -1).astype(np.float32)
    pcd[rs.rand(H, W) < 0.05] = np.nan
    return pcd, label


def integrate_frame(m, pcd, label, dev_inputs=False):
    if dev_inputs:
        pcd = torch.as_tensor(pcd).cuda()
        label = torch.as_tensor(label).cuda()
    for ins in list(range(1, 9)) + [0]:
        m.integrate(ins, label == ins, pcd)


def test_full_frame_properties(Mapping):
    pcd, label = full_frame(0)
    pitches = {i: 0.004 + 0.0005 * i for i in range(1, 9)}
    pitches[0] = 0.01

    def build(dev_inputs):
        m = Mapping()
        for ins in list(range(1, 9)) + [0]:
            m.initialize(ins, pitch=pitches[ins])
        integrate_frame(m, pcd, label, dev_inputs)
        pcd2, label2 = full_frame(1)
        integrate_frame(m, pcd2, label2, dev_inputs)
        return m
    a, b = build(False), build(True)
    # (1) thread order does not matter: two runs give the same map, cell for cell, bit for bit
    n = a.n_cells()
    assert n == b.n_cells() and n > 200000
    for ins in (1, 5, 0):
        ca, cb = a.cells(ins), b.cells(ins)
        assert ca == cb
        # (2) every log-odds value is one the sensor model can produce from <= 2 scans
        vals = set(np.float32(v) for v in ca.values())
        t = oc.OcTree(0.01)
        legal = set()
        for first in (t.prob_hit_log, t.prob_miss_log):
            legal.add(np.float32(first))
            for second in (t.prob_hit_log, t.prob_miss_log):
                legal.add(np.float32(np.float32(first) + second))
        assert vals <= legal, vals - legal
    # (3) a sample of rays checked against the oracle's DDA: every crossed cell is known, the end
    #     cell is occupied
    tree = oc.OcTree(pitches[3])
    cells3 = a.cells(3)
    ys, xs = np.nonzero((label == 3) & ~np.isnan(pcd).any(2))
    for j in range(0, len(ys), max(1, len(ys) // 40)):
        p = pcd[ys[j], xs[j]]
        for k in tree.computeRayKeys(np.zeros(3, np.float32), p):
            assert k in cells3
        assert cells3[tree.coords_to_key(p)] > 0
    # (4) one batched query == the single-target queries, and matches the numpy-returning API
    tids = list(range(1, 9))
    origins = []
    for t_ in tids:
        c = np.nanmedian(pcd[label == t_], axis=0)
        origins.append(c - 15.5 * pitches[t_])
    gt, gn, ge = a.get_target_grids_batch(tids, dimensions=(32, 32, 32), pitches=[pitches[t_] for t_ in tids],
                                          origins=origins)
    assert gt.is_cuda and gt.shape == (8, 32, 32, 32)
    for i, t_ in enumerate(tids):
        s = a.get_target_grids(t_, dimensions=(32, 32, 32), pitch=pitches[t_], origin=origins[i])
        for x, y in zip((gt[i], gn[i], ge[i]), s):
            assert np.array_equal(x.cpu().numpy(), y)
        assert (s[0] > 0).sum() > 50 and (s[2] > 0).sum() > 1000
    # (5) grids only hold values the reference's thresholds allow
    assert float(gt.min()) >= 0 and float(gt[gt > 0].min()) >= 0.5
    assert float(ge[ge > 0].min()) > 0.5                              # 1 - occ with occ < 0.5


def test_labelled_frame_equals_per_instance_scans_full_size(Mapping):
    pcd, label = full_frame(2)
    pitches = {i: 0.004 + 0.0005 * i for i in range(1, 9)}
    pitches[0] = 0.01
    a, b = Mapping(), Mapping()
    for m in (a, b):
        for ins in list(range(1, 9)) + [0]:
            m.initialize(ins, pitch=pitches[ins])
    integrate_frame(a, pcd, label)
    b.integrate_labels(torch.as_tensor(label).cuda(), torch.as_tensor(pcd).cuda())
    assert a.n_cells() == b.n_cells() > 100000
    ka, kb = a._cells.cpu(), b._cells.cpu()
    # same set of (key, log-odds) pairs, wherever the slots ended up
    ea = sorted(zip(ka[:, 0].tolist(), ka.view(torch.float32)[:, 2].tolist()))
    eb = sorted(zip(kb[:, 0].tolist(), kb.view(torch.float32)[:, 2].tolist()))
    assert ea == eb


def test_no_cpu_tensors(Mapping):
    m = Mapping()
    m.initialize(1, pitch=0.01)
    m.integrate(1, np.ones((2, 2), bool), np.ones((2, 2, 3), np.float32))     # numpy is uploaded
    with pytest.raises(ValueError, match="bool"):
        m.integrate(1, np.ones((2, 2), np.uint8), np.ones((2, 2, 3), np.float32))
