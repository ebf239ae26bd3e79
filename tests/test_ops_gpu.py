"""GPU parity tests: CUDA operators (through the C ABI) vs the oracle and vs the golden
vectors produced by the reference's own code.  Bit-exact for indices/counts/min-distances and
for the deterministic fp32 sums; stated tolerances elsewhere."""

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import transforms as otf
from oracle import voxel_ops as vo

pytestmark = pytest.mark.gpu
F32 = np.float32


def cu(x, dev, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x), device=dev)
    return t if dtype is None else t.to(dtype)


def F():
    import morefusion_b200
    return morefusion_b200.functions


# ------------------------------------------------------------------ a1
@pytest.mark.parametrize("case", ["unit32", "ties_oob"])
def test_average_voxelization_golden(cuda_device, case):
    g = golden("voxelization_" + case)
    dims = tuple(int(d) for d in g["dims"])
    values = cu(g["values"], cuda_device).requires_grad_(True)
    y, counts = F().average_voxelization_3d(
        values, cu(g["points"], cuda_device), cu(g["batch_indices"], cuda_device),
        batch_size=int(g["B"]), origin=g["origin"], pitch=g["pitch"], dimensions=dims,
        return_counts=True)
    assert y.dtype == torch.float32 and tuple(y.shape) == (int(g["B"]), values.shape[1]) + dims
    assert np.array_equal(counts.cpu().numpy(), g["ref_avg_counts_gpu"])   # bit-exact
    assert np.array_equal(y.detach().cpu().numpy(), g["ref_avg_matrix_gpu"])  # bit-exact
    gy = np.random.RandomState(int(g["gy_seed"])).uniform(-1, 1, tuple(y.shape)).astype(F32)
    y.backward(cu(gy, cuda_device))
    assert np.array_equal(values.grad.cpu().numpy(), g["ref_avg_gvalues_gpu"])


@pytest.mark.parametrize("shape", [
    dict(P=1024, C=4, B=1, D=32, sorted=True),      # BASELINE config 1
    dict(P=1024, C=144, B=1, D=32, sorted=True),
    dict(P=8000, C=144, B=8, D=32, sorted=True),    # model shape
    dict(P=5000, C=7, B=5, D=32, sorted=False),     # unsorted batch indices, odd C
    dict(P=3000, C=70, B=3, D=(7, 9, 11), sorted=False),  # non-cubic, V % 4 != 0, 2 chunks
    dict(P=0, C=3, B=2, D=8, sorted=True),          # empty input
    dict(P=20000, C=3, B=1, D=4, sorted=True),      # heavy collisions: ~300 points per voxel
    dict(P=60000, C=5, B=2, D=16, sorted=True),     # > 125 occupied voxels per segment (slot rounds), > 1024 keys per scan
    dict(P=40000, C=65, B=3, D=16, sorted=False),   # the same unsorted, 2 channel chunks
])
def test_average_voxelization_vs_oracle(cuda_device, shape):
    rs = np.random.RandomState(0)
    P, C, B = shape["P"], shape["C"], shape["B"]
    dims = shape["D"] if isinstance(shape["D"], tuple) else (shape["D"],) * 3
    pts = rs.uniform(-1.1, 1.1, (P, 3)).astype(F32)
    vals = rs.uniform(-1, 1, (P, C)).astype(F32)
    bi = rs.randint(0, B, P).astype(np.int32)
    if shape["sorted"]:
        bi = np.sort(bi)
    origin = np.array([-1, -1, -1], F32)
    pitch = F32(2.0 / max(dims))
    m, c = vo.average_voxelization_3d_fwd(
        vals, pts, bi, batch_size=B, origin=origin, pitch=pitch, dimensions=dims)
    v = cu(vals, cuda_device).requires_grad_(True)
    y, counts = F().average_voxelization_3d(
        v, cu(pts, cuda_device), cu(bi, cuda_device), batch_size=B, origin=origin, pitch=pitch,
        dimensions=dims, return_counts=True)
    assert np.array_equal(counts.cpu().numpy(), c)
    assert np.array_equal(y.detach().cpu().numpy(), m)
    if P:
        gy = rs.uniform(-1, 1, m.shape).astype(F32)
        y.backward(cu(gy, cuda_device))
        want = vo.average_voxelization_3d_bwd(gy, c, pts, bi, origin=origin, pitch=pitch,
                                              dimensions=dims)
        assert np.array_equal(v.grad.cpu().numpy(), want)


@pytest.mark.parametrize("seed", range(8))
def test_average_voxelization_oob_points_at_batch_boundaries(cuda_device, seed):
    """sorted batches whose boundary points are out of the grid, segment
    lengths not multiples of 8, on a workspace full of garbage (nothing in it may need
    initialising).  Seeds 1, 2 and 4 broke the round-1 leader kernel."""
    from morefusion_b200.functions.geometry import _util
    rs = np.random.RandomState(seed)
    P, C, B, D = 8000, 6, 8, 32
    pts = rs.uniform(-1.3, 1.3, (P, 3)).astype(F32)
    vals = rs.uniform(-1, 1, (P, C)).astype(F32)
    bi = np.sort(rs.randint(0, B, P)).astype(np.int32)
    if seed % 2:
        bi[bi == 3] = 4                               # an absent batch in the middle
    edges = np.flatnonzero(np.diff(bi)) + 1
    for e in edges:                                   # out-of-grid points either side of a boundary
        pts[max(e - rs.randint(1, 6), 0):e + rs.randint(1, 6)] = 5.0
    origin = np.array([-1, -1, -1], F32)
    pitch = F32(2.0 / D)
    m, c = vo.average_voxelization_3d_fwd(vals, pts, bi, batch_size=B, origin=origin, pitch=pitch,
                                          dimensions=(D, D, D))
    ws = _util.workspace(1 << 20, cuda_device)
    ws.view(torch.int32).fill_(int(rs.randint(-2 ** 31, 2 ** 31 - 1)))   # dirty workspace
    y, counts = F().average_voxelization_3d(
        cu(vals, cuda_device), cu(pts, cuda_device), cu(bi, cuda_device), batch_size=B,
        origin=origin, pitch=pitch, dimensions=(D, D, D), return_counts=True)
    assert np.array_equal(counts.cpu().numpy(), c)
    assert np.array_equal(y.cpu().numpy(), m)


def test_average_voxelization_properties_full_size(cuda_device):
    """Size-independent properties at the model shape: counts sum = in-bounds points,
    sum(matrix*counts) == sum(values of in-bounds points) per channel, empty voxels exactly 0."""
    torch.manual_seed(0)
    B, P, C, D = 8, 1000, 144, 32
    pts = torch.rand(B * P, 3, device=cuda_device) * 36 - 2
    vals = torch.randn(B * P, C, device=cuda_device)
    bi = torch.arange(B, device=cuda_device, dtype=torch.int32).repeat_interleave(P)
    y, counts = F().average_voxelization_3d(
        vals, pts, bi, batch_size=B, origin=(0, 0, 0), pitch=1.0, dimensions=(D, D, D),
        return_counts=True)
    idx = torch.round((pts - 0) / 1.0)
    inb = ((idx >= 0) & (idx < D)).all(1)
    assert int(counts.sum()) == int(inb.sum())
    tot = (y.double() * counts[:, None].double()).sum(dim=(0, 2, 3, 4))
    want = vals[inb].double().sum(0)
    torch.testing.assert_close(tot, want, rtol=1e-5, atol=1e-3)
    assert float(y[(counts == 0)[:, None].expand_as(y)].abs().max()) == 0.0


def test_average_voxelization_nan_raises(cuda_device):
    pts = torch.zeros(4, 3, device=cuda_device)
    pts[2, 1] = float("nan")
    with pytest.raises(ValueError, match="points include nan"):
        F().average_voxelization_3d(
            torch.zeros(4, 2, device=cuda_device), pts,
            torch.zeros(4, dtype=torch.int32, device=cuda_device),
            batch_size=1, origin=(0, 0, 0), pitch=1.0, dimensions=(2, 2, 2))


# ------------------------------------------------------------------ a6
@pytest.mark.parametrize("case", ["unit32", "ties_oob"])
def test_max_voxelization_golden(cuda_device, case):
    g = golden("voxelization_" + case)
    dims = tuple(int(d) for d in g["dims"])
    values = cu(g["values"], cuda_device).requires_grad_(True)
    y, ind = F().max_voxelization_3d(
        values, cu(g["points"], cuda_device), cu(g["batch_indices"], cuda_device),
        cu(g["intensities"], cuda_device), batch_size=int(g["B"]), origin=g["origin"],
        pitch=g["pitch"], dimensions=dims, return_indices=True)
    # the reference's serial-schedule result: first point wins exact intensity ties
    assert np.array_equal(ind.cpu().numpy(), g["ref_max_indices_gpu"])
    assert np.array_equal(y.detach().cpu().numpy(), g["ref_max_matrix_gpu"])
    gy = np.random.RandomState(int(g["gy_seed"])).uniform(-1, 1, tuple(y.shape)).astype(F32)
    y.backward(cu(gy, cuda_device))
    np.testing.assert_allclose(values.grad.cpu().numpy(), g["ref_max_gvalues_gpu"], rtol=0, atol=0)


# ------------------------------------------------------------------ a5
def test_interpolate_golden(cuda_device):
    g = golden("interpolate_16")
    vox = cu(g["voxelized"], cuda_device).requires_grad_(True)
    y = F().interpolate_voxel_grid(vox, cu(g["points"], cuda_device),
                                   cu(g["batch_indices"], cuda_device))
    assert np.array_equal(y.detach().cpu().numpy(), g["ref_values_gpu"])      # bit-exact
    y.backward(cu(g["gy"], cuda_device))
    # backward scatters with atomics: order-nondeterministic fp32 sums
    np.testing.assert_allclose(vox.grad.cpu().numpy(), g["ref_gvoxelized_gpu"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,C,D,P", [(8, 256, 16, 8000), (8, 512, 8, 8000), (2, 5, 6, 100)])
def test_interpolate_vs_oracle(cuda_device, B, C, D, P):
    rs = np.random.RandomState(1)
    vox = rs.uniform(-1, 1, (B, C, D, D, D)).astype(F32)
    pts = rs.uniform(-0.7, D - 0.3, (P, 3)).astype(F32)
    bi = rs.randint(0, B, P).astype(np.int32)
    want = vo.interpolate_voxel_grid_fwd(vox, pts, bi)
    y = F().interpolate_voxel_grid(cu(vox, cuda_device), cu(pts, cuda_device), cu(bi, cuda_device))
    assert np.array_equal(y.cpu().numpy(), want)
    # channels-last internal layout gives identical values
    from morefusion_b200.functions.geometry.interpolate_voxel_grid import InterpolateVoxelGrid
    vcl = cu(vox, cuda_device).permute(0, 2, 3, 4, 1).contiguous()
    y2 = InterpolateVoxelGrid.apply(vcl, cu(pts, cuda_device), cu(bi, cuda_device), True)
    assert np.array_equal(y2.cpu().numpy(), want)


@pytest.mark.parametrize("B,C,D,P", [(8, 256, 16, 8000), (8, 512, 8, 8000), (2, 5, 6, 100),
                                     (1, 12, 16, 9000), (3, 7, 20, 5000)])
def test_interpolate_backward_vs_oracle(cuda_device, B, C, D, P):
    """Gradient w.r.t. the grid at the model shapes (voxel-centric gather kernel), with unsorted
    batch indices, points partly outside the grid, a batch larger than the kernel's point cache
    (9000 points in one batch: multi-pass path) and a grid too large for the cell lists (20^3)."""
    from morefusion_b200.functions.geometry.interpolate_voxel_grid import InterpolateVoxelGrid
    rs = np.random.RandomState(4)
    pts = rs.uniform(-1.6, D + 0.5, (P, 3)).astype(F32)
    bi = rs.randint(0, B, P).astype(np.int32)
    gy = rs.uniform(-1, 1, (P, C)).astype(F32)
    want = vo.interpolate_voxel_grid_bwd(gy, pts, bi, (B, C, D, D, D))
    vox = torch.zeros((B, C, D, D, D), device=cuda_device, requires_grad=True)
    y = InterpolateVoxelGrid.apply(vox, cu(pts, cuda_device), cu(bi, cuda_device), False)
    y.backward(cu(gy, cuda_device))
    got = vox.grad.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-5)
    # the same through the forward: values at those points, bit-exact (multi-pass point cache)
    v = rs.uniform(-1, 1, (B, C, D, D, D)).astype(F32)
    yv = InterpolateVoxelGrid.apply(cu(v, cuda_device), cu(pts, cuda_device), cu(bi, cuda_device), False)
    assert np.array_equal(yv.cpu().numpy(), vo.interpolate_voxel_grid_fwd(v, pts, bi))
    # deterministic: a second run gives the same bits (no atomics on the gather path)
    vox2 = torch.zeros_like(vox, requires_grad=True)
    InterpolateVoxelGrid.apply(vox2, cu(pts, cuda_device), cu(bi, cuda_device), False).backward(cu(gy, cuda_device))
    if (D + 1) ** 3 <= 8192:
        assert torch.equal(vox.grad, vox2.grad)


# ------------------------------------------------------------------ a3 / a4
@pytest.mark.parametrize("case", ["main5", "ball16", "lattice_ties"])
def test_tdf_golden(cuda_device, case):
    g = golden("tdf_" + case)
    dims = tuple(int(d) for d in g["dims"])
    pts = cu(g["points"], cuda_device).requires_grad_(True)
    tdf, ind = F().truncated_distance_function(
        pts, pitch=g["pitch"], origin=g["origin"], dims=dims, truncation=g["truncation"],
        return_indices=True)
    assert np.array_equal(tdf.detach().cpu().numpy(), g["ref_tdf"])       # bit-exact
    assert np.array_equal(ind.cpu().numpy(), g["ref_indices"])            # deterministic winners
    tdf.backward(cu(g["gy"], cuda_device))
    np.testing.assert_allclose(pts.grad.cpu().numpy(), g["ref_gpoints"], rtol=1e-5, atol=1e-6)
    want = vo.truncated_distance_function_bwd(
        g["gy"], g["points"], g["ref_indices"], pitch=g["pitch"], origin=g["origin"], dims=dims)
    np.testing.assert_allclose(pts.grad.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    for off, tag in [(float(g["sdf_offset"]), "po"), (0.0, "po0")]:
        grid, surf, ins = F().pseudo_occupancy_voxelization(
            cu(g["points"], cuda_device), cu(g["sdf"], cuda_device), pitch=g["pitch"],
            origin=g["origin"], dims=dims, threshold=int(g["threshold"]), sdf_offset=off)
        if tag == "po":
            assert np.array_equal(grid.cpu().numpy(), g["ref_po_grid"], equal_nan=True)
        assert np.array_equal(surf.cpu().numpy(), g[f"ref_{tag}_surface"], equal_nan=True)
        assert np.array_equal(ins.cpu().numpy(), g[f"ref_{tag}_inside"], equal_nan=True)


def test_tdf_icc_shape_vs_oracle(cuda_device):
    rs = np.random.RandomState(2)
    P, D = 4000, 32
    pitch = F32(0.0063)
    pts = (rs.normal(0, 0.03, (P, 3)) + [0.1, 0.2, 0.7]).astype(F32)
    origin = (np.array([0.1, 0.2, 0.7], F32) - pitch * F32(15.5)).astype(F32)
    sdf = rs.uniform(-0.01, 0.03, P).astype(F32)
    r = vo.pseudo_occupancy_voxelization_fwd(pts, sdf, pitch=pitch, origin=origin, dims=(D,) * 3,
                                             threshold=2, sdf_offset=0.02)
    p = cu(pts, cuda_device).requires_grad_(True)
    grid, surf, ins = F().pseudo_occupancy_voxelization(
        p, cu(sdf, cuda_device), pitch=pitch, origin=origin, dims=(D,) * 3, threshold=2,
        sdf_offset=0.02)
    assert np.array_equal(grid.detach().cpu().numpy(), r["grid"])
    assert np.array_equal(surf.detach().cpu().numpy(), r["surface"])
    assert np.array_equal(ins.detach().cpu().numpy(), r["inside"])
    gs = rs.uniform(-1, 1, (D,) * 3).astype(F32)
    gi = rs.uniform(-1, 1, (D,) * 3).astype(F32)
    (surf * cu(gs, cuda_device) + ins * cu(gi, cuda_device)).sum().backward()
    d_grid = r["w_surface"] * gs + r["w_inside"] * gi
    want = vo.truncated_distance_function_bwd(
        (-(d_grid) / r["truncation"]).astype(F32), pts, r["indices"], pitch=pitch, origin=origin,
        dims=(D,) * 3)
    np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=2e-5, atol=1e-5)


# ------------------------------------------------------------------ a2
def test_occupancy_grid_3d(cuda_device):
    g = golden("occupancy_grid_3d")
    m = F().occupancy_grid_3d(cu(g["kat_points"], cuda_device), pitch=1, origin=(0, 0, 0),
                              dims=(5, 5, 5))
    nonzero = [[0, 0, 0], [0, 1, 0], [0, 0, 1], [4, 3, 4], [3, 4, 4], [4, 4, 4]]
    want = np.zeros((5, 5, 5), bool)
    want[tuple(zip(*nonzero))] = True
    assert np.array_equal(m.cpu().numpy() > 0, want)      # reference KAT, rtol=0 atol=0
    assert np.array_equal(m.cpu().numpy(), g["ref_kat"])
    dims = tuple(int(d) for d in g["dims"])
    pts = cu(g["points"], cuda_device).requires_grad_(True)
    m = F().occupancy_grid_3d(pts, pitch=g["pitch"], origin=g["origin"], dims=dims,
                              threshold=int(g["threshold"]))
    assert np.array_equal(m.detach().cpu().numpy(), g["ref_grid"])
    rs = np.random.RandomState(3)
    gm = rs.uniform(-1, 1, dims).astype(F32)
    m.backward(cu(gm, cuda_device))
    _, aux = vo.occupancy_grid_3d_fwd(g["points"], pitch=g["pitch"], origin=g["origin"], dims=dims,
                                      threshold=int(g["threshold"]))
    want = vo.occupancy_grid_3d_bwd(gm, aux, pitch=g["pitch"], threshold=int(g["threshold"]))
    np.testing.assert_allclose(pts.grad.cpu().numpy(), want, rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ a7
def test_transforms_golden(cuda_device):
    g = golden("transforms")
    f = F()
    q = cu(g["q"], cuda_device).requires_grad_(True)
    t = cu(g["t"], cuda_device).requires_grad_(True)
    R = f.quaternion_matrix(q)
    np.testing.assert_allclose(R.detach().cpu().numpy(), g["ref_R"], rtol=0, atol=2e-7)
    assert tuple(f.quaternion_matrix(q[0]).shape) == (4, 4)
    T = f.transformation_matrix(q, t)
    np.testing.assert_allclose(T.detach().cpu().numpy(), g["ref_T"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(
        f.transformation_matrix(q[1], t[1]).detach().cpu().numpy(), g["ref_T_single"], atol=2e-7)
    assert np.array_equal(
        f.compose_transform(cu(g["ref_R"][:, :3, :3], cuda_device), t).detach().cpu().numpy(),
        g["ref_compose"])
    assert np.array_equal(f.translation_matrix(t).detach().cpu().numpy(), g["ref_translation"])
    assert tuple(f.translation_matrix(t[0]).shape) == (4, 4)
    pts = cu(g["points"], cuda_device)
    X = f.transform_points(pts, cu(g["ref_T"], cuda_device))
    np.testing.assert_allclose(X.cpu().numpy(), g["ref_points_M"], rtol=1e-6, atol=1e-6)
    X1 = f.transform_points(pts, cu(g["ref_T"][2], cuda_device))
    np.testing.assert_allclose(X1.cpu().numpy(), g["ref_points_single"], rtol=1e-6, atol=1e-6)
    # backward of the whole chain vs the oracle's analytic reverse pass
    gR = cu(g["gR"], cuda_device)
    (R * gR).sum().backward()
    _, aux = otf.quaternion_matrix_fwd(g["q"])
    np.testing.assert_allclose(q.grad.cpu().numpy(), otf.quaternion_matrix_bwd(g["gR"], aux),
                               rtol=1e-4, atol=1e-5)


def test_transform_points_backward(cuda_device):
    rs = np.random.RandomState(4)
    P, M = 300, 5
    pts = torch.tensor(rs.uniform(-1, 1, (P, 3)).astype(F32), device=cuda_device, requires_grad=True)
    T = torch.tensor(rs.uniform(-1, 1, (M, 4, 4)).astype(F32), device=cuda_device, requires_grad=True)
    go = torch.tensor(rs.uniform(-1, 1, (M, P, 3)).astype(F32), device=cuda_device)
    out = F().transform_points(pts, T)
    out.backward(go)
    p2 = pts.detach().clone().requires_grad_(True)
    T2 = T.detach().clone().requires_grad_(True)
    ph = torch.cat([p2, torch.ones(P, 1, device=cuda_device)], 1)
    ref = torch.matmul(T2, ph.T).transpose(1, 2)[:, :, :3]
    ref.backward(go)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(pts.grad, p2.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(T.grad[:, :3], T2.grad[:, :3], rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------ a12
def test_average_distance_golden_and_grad(cuda_device):
    from oracle import loss as oloss
    g = golden("average_distance")
    f = F()
    pts = cu(g["points"], cuda_device)
    for sym, key in ((False, "ref_add"), (True, "ref_add_s")):
        T1 = cu(g["T1"], cuda_device).requires_grad_(True)
        T2 = cu(g["T2"], cuda_device).requires_grad_(True)
        out = f.average_distance(pts, T1, T2, symmetric=sym)
        np.testing.assert_allclose(out.detach().cpu().numpy(), g[key], rtol=1e-6, atol=1e-7)
        # rows 2.. have non-zero distances everywhere (row 0 is the identical pose: sqrt'(0) = inf)
        go = np.zeros(5, F32)
        go[2:] = np.array([0.5, -1.0, 2.0], F32)
        want = oloss.average_distance(g["points"], g["T1"], g["T2"][2:], symmetric=sym,
                                      return_grads=True, gout=go[2:])
        out2 = f.average_distance(pts, T1, T2[2:], symmetric=sym)
        out2.backward(cu(go[2:], cuda_device))
        np.testing.assert_allclose(T2.grad[2:].cpu().numpy(), want[1], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(T1.grad.cpu().numpy(), want[2], rtol=1e-4, atol=1e-4)


def test_average_distance_training_shape(cuda_device):
    """Training-like shape (model.py:417,429: 500 CAD points per object, one pose per point),
    symmetric: the matrix-free NN equals a brute-force argmin over exact fp32 squared distances
    (torch.cdist's |a|^2+|b|^2-2ab expansion flips near-ties); gradients equal torch autograd."""
    torch.manual_seed(0)
    P, M = 500, 200
    pts = torch.rand(P, 3, device=cuda_device) * 0.2 - 0.1
    def rand_T(n):
        q = torch.randn(n, 4, device=cuda_device)
        T = F().transformation_matrix(q, torch.rand(n, 3, device=cuda_device) * 0.02)
        return T
    Tt = rand_T(1)[0].detach().requires_grad_(True)
    Tp = (rand_T(M).detach() * 1.0).requires_grad_(True)
    out = F().average_distance(pts, Tt, Tp, symmetric=True)
    out.sum().backward()
    Tt2, Tp2 = Tt.detach().clone().requires_grad_(True), Tp.detach().clone().requires_grad_(True)
    a = pts @ Tt2[:3, :3].T + Tt2[:3, 3]
    b = torch.einsum("mij,pj->mpi", Tp2[:, :3, :3], pts) + Tp2[:, None, :3, 3]
    diff = b.detach()[:, :, None, :] - a.detach()[None, None, :, :]
    idx = ((diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]).argmin(2)
    ref = (a[idx] - b).norm(dim=2).mean(1)
    ref.sum().backward()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(Tp.grad[:, :3], Tp2.grad[:, :3], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(Tt.grad[:3], Tt2.grad[:3], rtol=1e-3, atol=1e-2)
