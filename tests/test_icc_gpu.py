"""GPU parity: fused ICC kernel vs (1) the reference's own forward loss (golden, produced by
reference code), (2) the oracle's loss / analytic gradient, (3) the same graph composed from the
individual CUDA operators under torch autograd, (4) the oracle's Chainer-Adam refinement loop:
final R/t within 1e-4 abs (BASELINE north_star tolerance)."""

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import icc as oicc
from oracle import transforms as otf

pytestmark = pytest.mark.gpu
F32 = np.float32


def _inputs(g, dev):
    n = int(g["n_objects"])
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    return dict(points=[t(g[f"points_{i}"]) for i in range(n)],
                sdf=[t(g[f"sdf_{i}"]) for i in range(n)], pitch=t(g["pitch"]),
                origin=t(g["origin"]), grid_target=t(g["grid_target"]),
                grid_nontarget_empty=t(g["grid_nontarget_empty"]))


def _np_inputs(g):
    n = int(g["n_objects"])
    return dict(points=[g[f"points_{i}"] for i in range(n)], sdf=[g[f"sdf_{i}"] for i in range(n)],
                pitch=g["pitch"], origin=g["origin"], grid_target=g["grid_target"],
                grid_nontarget_empty=g["grid_nontarget_empty"])


def _link(g, dev):
    from morefusion_b200.contrib import IterativeCollisionCheckLink
    link = IterativeCollisionCheckLink(
        g["transform_init"], voxel_dim=int(g["voxel_dim"]), voxel_threshold=int(g["voxel_threshold"]),
        sdf_offset=float(g["sdf_offset"])).to(dev)
    return link


@pytest.mark.parametrize("case", ["contact3", "isolated2", "single1"])
def test_icc_forward_backward(cuda_device, case):
    g = golden("icc_forward_" + case)
    link = _link(g, cuda_device)
    np.testing.assert_allclose(link.quaternion.detach().cpu().numpy(), g["ref_quaternion"], atol=1e-6)
    inp = _inputs(g, cuda_device)
    loss = link(inp["points"], inp["sdf"], inp["pitch"], inp["origin"], inp["grid_target"],
                inp["grid_nontarget_empty"])
    # (1) the reference's own forward value
    np.testing.assert_allclose(float(loss), float(g["ref_loss"]), rtol=5e-5, atol=5e-6)
    loss.backward()
    # (2) oracle loss + analytic gradient
    r = oicc.icc_forward_backward(
        link.quaternion.detach().cpu().numpy(), link.translation.detach().cpu().numpy(),
        voxel_dim=int(g["voxel_dim"]), voxel_threshold=int(g["voxel_threshold"]),
        sdf_offset=float(g["sdf_offset"]), **_np_inputs(g))
    np.testing.assert_allclose(float(loss), r["loss"], rtol=5e-5, atol=5e-6)
    gq, gt = link.quaternion.grad.cpu().numpy(), link.translation.grad.cpu().numpy()
    np.testing.assert_allclose(gt, r["gt"], rtol=2e-3, atol=2e-3 * np.abs(r["gt"]).max())
    np.testing.assert_allclose(gq, r["gq"], rtol=2e-3, atol=2e-3 * np.abs(r["gq"]).max())


def test_icc_matches_composed_operators(cuda_device):
    """The fused kernel == the reference's forward graph built from the individual CUDA
    operators (iterative_collision_check_link.py:31-99) under torch autograd."""
    import morefusion_b200.functions as F
    g = golden("icc_forward_contact3")
    link = _link(g, cuda_device)
    inp = _inputs(g, cuda_device)
    loss = link(inp["points"], inp["sdf"], inp["pitch"], inp["origin"], inp["grid_target"],
                inp["grid_nontarget_empty"])
    loss.backward()
    q = link.quaternion.detach().clone().requires_grad_(True)
    t = link.translation.detach().clone().requires_grad_(True)
    D = int(g["voxel_dim"])
    T = F.transformation_matrix(q, t)
    pts = [F.transform_points(p, T[i]) for i, p in enumerate(inp["points"])]
    gs, gi, gne = [], [], []
    N = len(pts)
    for i in range(N):
        _, s, ins = F.pseudo_occupancy_voxelization(
            pts[i], inp["sdf"][i], pitch=inp["pitch"][i], origin=inp["origin"][i], dims=(D,) * 3,
            threshold=2, sdf_offset=0.02)
        gs.append(s)
        gi.append(ins)
        e = inp["grid_nontarget_empty"][i]
        po = torch.cat([p for j, p in enumerate(pts) if j != i], 0)
        so = torch.cat([p for j, p in enumerate(inp["sdf"]) if j != i], 0)
        _, _, other = F.pseudo_occupancy_voxelization(
            po, so, pitch=inp["pitch"][i], origin=inp["origin"][i], dims=(D,) * 3, threshold=2)
        if not torch.isnan(other).any():
            e = torch.where(e >= other, e, other)       # F.maximum, first arg on ties
        gne.append(e)
    gs, gi, gne = torch.stack(gs), torch.stack(gi), torch.stack(gne)
    reward = (gs * inp["grid_target"]).sum() / inp["grid_target"].sum()
    penalty = (gi * gne).sum() / gi.sum()
    loss2 = penalty - reward
    loss2.backward()
    torch.testing.assert_close(loss.detach(), loss2.detach(), rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(link.translation.grad, t.grad, rtol=2e-3, atol=2e-3 * float(t.grad.abs().max()))
    torch.testing.assert_close(link.quaternion.grad, q.grad, rtol=2e-3, atol=2e-3 * float(q.grad.abs().max()))


@pytest.mark.parametrize("case,n_iter", [("contact3", 30), ("isolated2", 5), ("single1", 5)])
def test_icc_refine_vs_oracle(cuda_device, case, n_iter):
    g = golden("icc_forward_" + case)
    link = _link(g, cuda_device)
    inp = _inputs(g, cuda_device)
    hist = link.refine(inp["points"], inp["sdf"], inp["pitch"], inp["origin"], inp["grid_target"],
                       inp["grid_nontarget_empty"], n_iter=n_iter)
    q_ref, t_ref, h_ref = oicc.icc_refine(
        g["transform_init"], n_iter=n_iter, voxel_dim=int(g["voxel_dim"]),
        sdf_offset=float(g["sdf_offset"]), return_history=True, **_np_inputs(g))
    assert hist.shape == (n_iter,)
    np.testing.assert_allclose(hist.cpu().numpy()[:3], h_ref[:3], rtol=1e-4, atol=1e-5)
    # pose R/t within 1e-4 abs
    T = otf.transformation_matrix(link.quaternion.detach().cpu().numpy(),
                                  link.translation.detach().cpu().numpy())
    T_ref = otf.transformation_matrix(q_ref, t_ref)
    np.testing.assert_allclose(T[:, :3, 3], T_ref[:, :3, 3], rtol=0, atol=1e-4)
    # The golden objects are balls.  In contact their rotation is constrained by the neighbours'
    # collision term; an ISOLATED ball's rotation is unobservable (gradient = round-off), and
    # Adam's m/sqrt(v) normalisation amplifies round-off to +-alpha steps -- in the reference
    # (racy fp32 atomics) as much as here.  Rotation parity is asserted where it is defined.
    rot_atol = 1e-4 if case == "contact3" else 2e-2
    np.testing.assert_allclose(T[:, :3, :3], T_ref[:, :3, :3], rtol=0, atol=rot_atol)


def test_icc_full_size_scene_and_batch(cuda_device):
    """8 objects, 32^3 grids, ~4-5k SDF points each (BASELINE config 4 shape).

    (a) open loop: loss and gradients of the fused kernel along the ORACLE's trajectory agree
        to fp32 round-off at every iteration (<= 2e-5 of the per-object gradient magnitude);
    (b) closed loop (fused Chainer-Adam, 5 iterations): translation within 1e-4, quaternion
        within 1e-3 -- Adam divides each COMPONENT by its own sqrt(v), so components whose
        gradient is ~1e-4 of the object's largest one inherit a ~1e-2 relative round-off and
        move by a few % of alpha differently; the reference's racy fp32 atomics behave the same;
    (c) the same scene three times in one batched launch reproduces the single-scene run."""
    from morefusion_b200 import synthetic
    from morefusion_b200.contrib import IterativeCollisionCheckLink
    from morefusion_b200.contrib.iterative_collision_check_link import ICCBatch
    # boxes with three distinct extents: every pose degree of freedom is observable
    sc = synthetic.make_icc_scene(N=8, seed=3, kinds=("box",))
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=cuda_device)   # noqa: E731
    pts, sdf = [t(p) for p in sc["points"]], [t(s) for s in sc["sdf"]]
    args = (pts, sdf, t(sc["pitch"]), t(sc["origin"]), t(sc["grid_target"]), t(sc["grid_nontarget_empty"]))
    np_args = (sc["points"], sc["sdf"], sc["pitch"], sc["origin"], sc["grid_target"],
               sc["grid_nontarget_empty"])
    # (a)
    link = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(cuda_device)
    q = np.stack([otf.quaternion_from_matrix(T) for T in sc["transform_init"]]).astype(F32)
    tr = sc["transform_init"][:, :3, 3].astype(F32).copy()
    oq, ot = oicc.ChainerAdam(q.shape, 0.01), oicc.ChainerAdam(tr.shape, 0.001)
    for _ in range(3):
        r = oicc.icc_forward_backward(q, tr, *np_args, sdf_offset=0.02)
        with torch.no_grad():
            link.quaternion.copy_(t(q))
            link.translation.copy_(t(tr))
        link.zero_grad()
        loss = link(*args)
        loss.backward()
        # fp64 sums of identical fp32 terms on both sides: equal up to the final rounding
        np.testing.assert_allclose(float(loss.detach()), r["loss"], rtol=3e-7, atol=0)
        gq, gt = link.quaternion.grad.cpu().numpy(), link.translation.grad.cpu().numpy()
        np.testing.assert_allclose(gt, r["gt"], rtol=3e-7, atol=0)
        assert (np.abs(gq - r["gq"]) / np.abs(r["gq"]).max(1, keepdims=True)).max() < 1e-6
        oq.update(q, r["gq"])
        ot.update(tr, r["gt"])
    # (b)
    link = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(cuda_device)
    n_iter = 5
    hist = link.refine(*args, n_iter=n_iter)
    q_ref, t_ref, h_ref = oicc.icc_refine(sc["transform_init"], *np_args, n_iter=n_iter,
                                          sdf_offset=0.02, return_history=True)
    np.testing.assert_allclose(hist.cpu().numpy(), h_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(link.translation.detach().cpu().numpy(), t_ref, rtol=0, atol=1e-6)
    np.testing.assert_allclose(link.quaternion.detach().cpu().numpy(), q_ref, rtol=0, atol=1e-6)
    # (c)
    batch = ICCBatch([sc, sc, sc], sdf_offset=0.02, device=cuda_device)
    hb = batch.refine(n_iter=n_iter)
    assert hb.shape == (3, n_iter)
    for s in range(3):
        np.testing.assert_allclose(hb[s].cpu().numpy(), hist.cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(batch.translation[8 * s:8 * s + 8].cpu().numpy(),
                                   link.translation.detach().cpu().numpy(), atol=1e-4)


# ------------------------------------------------------------------ closed loop, BASELINE config 4
@pytest.mark.parametrize("name", ["ref3", "seed3", "seed4"])
def test_icc_closed_loop_100_iterations(cuda_device, name):
    """north_star: pose R/t within 1e-4 abs after the full driver loop
    (check_iterative_collision_check_link.py:44-79: 100 iterations, Adam 0.01 / 0.001,
    sdf_offset 0.02) on the two 8-object scenes of BASELINE config 4 and on the reference's own
    committed 3-object scene.  The oracle trajectories were computed once by
    oracle/ref_harness/gen_icc_closed_loop.py (100 s of NumPy each) and are stored in
    tests/golden; the inputs are checked against the stored checksum."""
    from morefusion_b200.contrib import IterativeCollisionCheckLink
    from oracle.ref_harness import gen_icc_closed_loop as gen
    g = golden("icc_closed_loop_" + name)
    if name == "ref3":
        off = np.r_[0, np.cumsum(g["sizes"])]
        sc = dict(points=[g["points"][off[i]:off[i + 1]] for i in range(3)],
                  sdf=[g["sdf"][off[i]:off[i + 1]] for i in range(3)], pitch=g["pitch"],
                  origin=g["origin"], grid_target=g["grid_target"].astype(F32),
                  grid_nontarget_empty=g["grid_nontarget_empty"].astype(F32),
                  transform_init=g["transform_init"])
    else:
        sc = gen.scene(name)
    assert gen.inputs_checksum(sc) == str(g["inputs_sha1"]), "inputs differ from the fixture's"
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=cuda_device)   # noqa: E731
    link = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(cuda_device)
    # start from the fixture's initial parameters: quaternion_from_matrix (LAPACK eigh) is not
    # bit-reproducible across CPUs and the loop amplifies 1 ulp to O(alpha) within 100 steps
    assert np.abs(link.quaternion.detach().cpu().numpy() - g["q0"]).max() < 1e-6
    with torch.no_grad():
        link.quaternion.copy_(t(g["q0"]))
        link.translation.copy_(t(g["t0"]))
    n_iter = int(g["n_iter"])
    hist = link.refine([t(p) for p in sc["points"]], [t(x) for x in sc["sdf"]], t(sc["pitch"]),
                       t(sc["origin"]), t(sc["grid_target"]), t(sc["grid_nontarget_empty"]),
                       n_iter=n_iter)
    T = otf.transformation_matrix(link.quaternion.detach().cpu().numpy(),
                                  link.translation.detach().cpu().numpy())
    T_ref = otf.transformation_matrix(g["q"], g["t"])
    dt = np.abs(T[:, :3, 3] - T_ref[:, :3, 3]).max()
    dR = np.abs(T[:, :3, :3] - T_ref[:, :3, :3]).max()
    dl = np.abs(hist.cpu().numpy() - g["loss"]).max()
    print(f"closed loop {name}: max|dt|={dt:.3g} max|dR|={dR:.3g} max|dloss|={dl:.3g}")
    assert dt <= 1e-4 and dR <= 1e-4, (dt, dR)
    np.testing.assert_allclose(hist.cpu().numpy(), g["loss"], rtol=1e-4, atol=1e-5)
