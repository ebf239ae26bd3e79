"""Pin the oracle: oracle == outputs of the reference's own code (tests/golden),
bit-exact wherever the operation is integer / order-deterministic.

Golden arrays ``ref_*`` were produced by oracle/ref_harness/gen_golden.py by
executing reference code (see that file).  Runs on CPU, no GPU needed.
"""

import numpy as np
import pytest

from conftest import golden
from oracle import icc as oicc
from oracle import transforms as otf
from oracle import voxel_ops as vo

F32 = np.float32


@pytest.mark.parametrize("case", ["unit32", "ties_oob"])
@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_average_voxelization(case, mode):
    g = golden("voxelization_" + case)
    dims = tuple(int(d) for d in g["dims"])
    kw = dict(origin=g["origin"], pitch=g["pitch"], dimensions=dims,
              numpy_semantics=(mode == "cpu"))
    m, c = vo.average_voxelization_3d_fwd(
        g["values"], g["points"], g["batch_indices"], batch_size=int(g["B"]), **kw)
    assert np.array_equal(c, g[f"ref_avg_counts_{mode}"])          # counts: bit-exact
    assert np.array_equal(m, g[f"ref_avg_matrix_{mode}"])          # serial fp32 sums: bit-exact
    gy = np.random.RandomState(int(g["gy_seed"])).uniform(-1, 1, m.shape).astype(F32)
    gv = vo.average_voxelization_3d_bwd(gy, c, g["points"], g["batch_indices"], **kw)
    assert np.array_equal(gv, g[f"ref_avg_gvalues_{mode}"])


def test_average_voxelization_rounding_modes_differ():
    # the ties case must actually exercise half-away vs half-even
    g = golden("voxelization_ties_oob")
    assert not np.array_equal(g["ref_avg_counts_cpu"], g["ref_avg_counts_gpu"])


@pytest.mark.parametrize("case", ["unit32", "ties_oob"])
@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_max_voxelization(case, mode):
    g = golden("voxelization_" + case)
    dims = tuple(int(d) for d in g["dims"])
    m, ind = vo.max_voxelization_3d_fwd(
        g["values"], g["points"], g["batch_indices"], g["intensities"],
        batch_size=int(g["B"]), origin=g["origin"], pitch=g["pitch"], dimensions=dims,
        numpy_semantics=(mode == "cpu"))
    assert np.array_equal(ind, g[f"ref_max_indices_{mode}"])
    assert np.array_equal(m, g[f"ref_max_matrix_{mode}"])
    gy = np.random.RandomState(int(g["gy_seed"])).uniform(-1, 1, m.shape).astype(F32)
    gv = vo.max_voxelization_3d_bwd(gy, ind, g["points"].shape[0])
    np.testing.assert_allclose(gv, g[f"ref_max_gvalues_{mode}"], rtol=1e-6, atol=1e-6)


def test_voxelization_errors():
    with pytest.raises(ValueError):
        vo.average_voxelization_3d_fwd(
            np.zeros((1, 1), F32), np.full((1, 3), np.nan, F32), np.zeros(1, np.int32),
            batch_size=1, origin=(0, 0, 0), pitch=1.0, dimensions=(2, 2, 2))
    with pytest.raises(ValueError):
        vo.average_voxelization_3d_fwd(
            np.zeros((1, 1), F32), np.zeros((1, 3), F32), np.zeros(1, np.int32),
            batch_size=1, origin=(0, 0, 0), pitch=1.0, dimensions=[2, 2, 2])


@pytest.mark.parametrize("mode", ["cpu", "gpu"])
def test_interpolate(mode):
    g = golden("interpolate_16")
    y = vo.interpolate_voxel_grid_fwd(
        g["voxelized"], g["points"], g["batch_indices"], numpy_semantics=(mode == "cpu"))
    assert np.array_equal(y, g[f"ref_values_{mode}"])
    if mode == "gpu":
        gv = vo.interpolate_voxel_grid_bwd(
            g["gy"], g["points"], g["batch_indices"], g["voxelized"].shape)
        assert np.array_equal(gv, g["ref_gvoxelized_gpu"])


@pytest.mark.parametrize("case", ["main5", "ball16", "lattice_ties"])
def test_tdf_and_pseudo_occupancy(case):
    g = golden("tdf_" + case)
    dims = tuple(int(d) for d in g["dims"])
    kw = dict(pitch=g["pitch"], origin=g["origin"], dims=dims)
    assert vo.tdf_ksize(g["pitch"], g["truncation"]) == int(g["ksize"])
    tdf, ind = vo.truncated_distance_function_fwd(g["points"], truncation=g["truncation"], **kw)
    assert np.array_equal(tdf, g["ref_tdf"])                 # min distances: bit-exact
    assert np.array_equal(ind, g["ref_indices"])             # winners: serial-schedule tie-break
    gp = vo.truncated_distance_function_bwd(g["gy"], g["points"], ind, **kw)
    np.testing.assert_allclose(gp, g["ref_gpoints"], rtol=1e-5, atol=1e-6)
    for off, tag in [(float(g["sdf_offset"]), "po"), (0.0, "po0")]:
        with np.errstate(invalid="ignore"):
            r = vo.pseudo_occupancy_voxelization_fwd(
                g["points"], g["sdf"], threshold=int(g["threshold"]), sdf_offset=off, **kw)
        if tag == "po":
            assert np.array_equal(r["grid"], g["ref_po_grid"], equal_nan=True)
        assert np.array_equal(r["surface"], g[f"ref_{tag}_surface"], equal_nan=True)
        assert np.array_equal(r["inside"], g[f"ref_{tag}_inside"], equal_nan=True)


def test_occupancy_grid_3d():
    g = golden("occupancy_grid_3d")
    # reference known-answer test: tests/functions_tests/geometry_tests/test_occupancy_grid_3d.py:24-38
    m, _ = vo.occupancy_grid_3d_fwd(g["kat_points"], pitch=1, origin=(0, 0, 0), dims=(5, 5, 5))
    nonzero = [[0, 0, 0], [0, 1, 0], [0, 0, 1], [4, 3, 4], [3, 4, 4], [4, 4, 4]]
    want = np.zeros((5, 5, 5), bool)
    want[tuple(zip(*nonzero))] = True
    assert np.array_equal(m > 0, want)
    assert np.array_equal(m, g["ref_kat"])
    dims = tuple(int(d) for d in g["dims"])
    m, aux = vo.occupancy_grid_3d_fwd(
        g["points"], pitch=g["pitch"], origin=g["origin"], dims=dims, threshold=int(g["threshold"]))
    assert np.array_equal(m, g["ref_grid"])
    # reference OccupancyGrid3D.backward (occupancy_grid_3d.py:56-74): gd -> gpoints
    gp = np.stack([(-g[f"gd{k}"] / g["pitch"]).sum(axis=(0, 1, 2)) for k in range(3)], 1)
    np.testing.assert_allclose(gp, g["ref_gpoints_from_gd"], rtol=1e-5, atol=1e-4)


def test_occupancy_grid_3d_backward_numeric():
    rs = np.random.RandomState(0)
    pts = rs.uniform(0.5, 3.5, (6, 3)).astype(F32)
    gm = rs.uniform(-1, 1, (5, 5, 5)).astype(F32)
    kw = dict(pitch=F32(1.0), origin=np.zeros(3, F32), dims=(5, 5, 5), threshold=2)
    m, aux = vo.occupancy_grid_3d_fwd(pts, **kw)
    gp = vo.occupancy_grid_3d_bwd(gm, aux, pitch=kw["pitch"], threshold=2)
    num = np.zeros_like(pts, dtype=np.float64)
    eps = 1e-3
    for i in range(pts.shape[0]):
        for k in range(3):
            a, b = pts.copy(), pts.copy()
            a[i, k] += eps
            b[i, k] -= eps
            fa = (vo.occupancy_grid_3d_fwd(a, **kw)[0].astype(np.float64) * gm).sum()
            fb = (vo.occupancy_grid_3d_fwd(b, **kw)[0].astype(np.float64) * gm).sum()
            num[i, k] = (fa - fb) / (2 * eps)
    np.testing.assert_allclose(gp, num, rtol=5e-2, atol=5e-3)


def test_transforms():
    g = golden("transforms")
    R, aux = otf.quaternion_matrix_fwd(g["q"])
    np.testing.assert_allclose(R, g["ref_R"], rtol=0, atol=2e-7)
    R1, _ = otf.quaternion_matrix_fwd(g["q"][0])
    np.testing.assert_allclose(R1, g["ref_R_single"], rtol=0, atol=2e-7)
    # rotation matrices: orthonormal, det +1 (vs trimesh KAT of test_quaternion_matrix.py:29-37)
    Rm = R[:, :3, :3].astype(np.float64)
    np.testing.assert_allclose(Rm @ Rm.transpose(0, 2, 1), np.tile(np.eye(3), (7, 1, 1)), atol=1e-6)
    np.testing.assert_allclose(np.linalg.det(Rm), 1, atol=1e-6)
    assert np.array_equal(otf.compose_transform(g["ref_R"][:, :3, :3], g["t"]), g["ref_compose"])
    assert np.array_equal(otf.translation_matrix(g["t"]), g["ref_translation"])
    np.testing.assert_allclose(otf.transformation_matrix(g["q"], g["t"]), g["ref_T"], atol=2e-7)
    np.testing.assert_allclose(
        otf.transformation_matrix(g["q"][1], g["t"][1]), g["ref_T_single"], atol=2e-7)
    np.testing.assert_allclose(
        otf.transform_points(g["points"], g["ref_T"]), g["ref_points_M"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(
        otf.transform_points(g["points"], g["ref_T"][2]), g["ref_points_single"], rtol=1e-6, atol=1e-6)


def test_quaternion_matrix_backward():
    g = golden("transforms")
    # table backward: oracle's first stage == reference QuaternionMatrix.backward
    Q = g["Q"]
    q = g["q"]
    R, aux = otf.quaternion_matrix_fwd(q)
    gR = g["gR"]
    gq = otf.quaternion_matrix_bwd(gR, aux)
    # numeric check of the full chain in float64
    def f(qq):
        qq = qq.astype(np.float64)
        n = (qq * qq).sum(1, keepdims=True)
        s = qq * np.sqrt(2.0 / n)
        Qm = s[:, :, None] * s[:, None, :]
        Rr = np.tile(np.eye(4)[None], (qq.shape[0], 1, 1))
        Rr[:, 0, 0] = 1 - Qm[:, 2, 2] - Qm[:, 3, 3]; Rr[:, 0, 1] = Qm[:, 1, 2] - Qm[:, 3, 0]
        Rr[:, 0, 2] = Qm[:, 1, 3] + Qm[:, 2, 0]; Rr[:, 1, 0] = Qm[:, 1, 2] + Qm[:, 3, 0]
        Rr[:, 1, 1] = 1 - Qm[:, 1, 1] - Qm[:, 3, 3]; Rr[:, 1, 2] = Qm[:, 2, 3] - Qm[:, 1, 0]
        Rr[:, 2, 0] = Qm[:, 1, 3] - Qm[:, 2, 0]; Rr[:, 2, 1] = Qm[:, 2, 3] + Qm[:, 1, 0]
        Rr[:, 2, 2] = 1 - Qm[:, 1, 1] - Qm[:, 2, 2]
        return (Rr * gR).sum()
    num = np.zeros_like(q, dtype=np.float64)
    eps = 1e-5
    for i in range(q.shape[0]):
        for k in range(4):
            a, b = q.astype(np.float64), q.astype(np.float64)
            a[i, k] += eps; b[i, k] -= eps
            num[i, k] = (f(a) - f(b)) / (2 * eps)
    np.testing.assert_allclose(gq, num, rtol=2e-4, atol=2e-5)
    # and the reference's backward table itself
    gQ = np.zeros_like(gR)
    gQ[:, 1, 0] = -gR[:, 1, 2] + gR[:, 2, 1]; gQ[:, 1, 1] = -gR[:, 1, 1] - gR[:, 2, 2]
    gQ[:, 1, 2] = gR[:, 0, 1] + gR[:, 1, 0]; gQ[:, 1, 3] = gR[:, 0, 2] + gR[:, 2, 0]
    gQ[:, 2, 0] = gR[:, 0, 2] - gR[:, 2, 0]; gQ[:, 2, 2] = -gR[:, 0, 0] - gR[:, 2, 2]
    gQ[:, 2, 3] = gR[:, 1, 2] + gR[:, 2, 1]; gQ[:, 3, 0] = -gR[:, 0, 1] + gR[:, 1, 0]
    gQ[:, 3, 3] = -gR[:, 0, 0] - gR[:, 1, 1]
    assert np.array_equal(gQ, g["ref_table_bwd"])


def _icc_inputs(g):
    n = int(g["n_objects"])
    return dict(
        points=[g[f"points_{i}"] for i in range(n)], sdf=[g[f"sdf_{i}"] for i in range(n)],
        pitch=g["pitch"], origin=g["origin"], grid_target=g["grid_target"],
        grid_nontarget_empty=g["grid_nontarget_empty"])


@pytest.mark.parametrize("case", ["contact3", "isolated2", "single1"])
def test_icc_forward_loss(case):
    g = golden("icc_forward_" + case)
    q = np.stack([otf.quaternion_from_matrix(T) for T in g["transform_init"]]).astype(F32)
    np.testing.assert_allclose(q, g["ref_quaternion"], atol=1e-7)
    kw = _icc_inputs(g)
    r = oicc.icc_forward_backward(
        g["ref_quaternion"], g["ref_translation"], voxel_dim=int(g["voxel_dim"]),
        voxel_threshold=int(g["voxel_threshold"]), sdf_offset=float(g["sdf_offset"]),
        need_grad=False, **kw)
    assert np.isfinite(r["loss"])
    np.testing.assert_allclose(r["loss"], g["ref_loss"], rtol=2e-5, atol=2e-6)
    if case == "isolated2":
        assert r["other_used"] == [False, False]
    if case == "contact3":
        assert all(r["other_used"])


def test_icc_gradient_autograd():
    """Oracle's hand-derived reverse pass vs torch autograd over the same graph
    (winners and weights frozen, exactly what the chainer graph differentiates:
    truncated_distance_function.py:196-213 builds the weights from raw arrays)."""
    import torch
    g = golden("icc_forward_contact3")
    kw = _icc_inputs(g)
    q0, t0 = g["ref_quaternion"].copy(), g["ref_translation"].copy()
    D = int(g["voxel_dim"])
    r = oicc.icc_forward_backward(q0, t0, voxel_dim=D, voxel_threshold=2, sdf_offset=0.02, **kw)
    N = q0.shape[0]
    dd = torch.float64
    q = torch.tensor(q0, dtype=dd, requires_grad=True)
    t = torch.tensor(t0, dtype=dd, requires_grad=True)
    n = (q * q).sum(1, keepdim=True)
    s = q * torch.sqrt(2.0 / n)
    Q = s[:, :, None] * s[:, None, :]
    R = torch.stack([
        torch.stack([1 - Q[:, 2, 2] - Q[:, 3, 3], Q[:, 1, 2] - Q[:, 3, 0], Q[:, 1, 3] + Q[:, 2, 0]], 1),
        torch.stack([Q[:, 1, 2] + Q[:, 3, 0], 1 - Q[:, 1, 1] - Q[:, 3, 3], Q[:, 2, 3] - Q[:, 1, 0]], 1),
        torch.stack([Q[:, 1, 3] - Q[:, 2, 0], Q[:, 2, 3] + Q[:, 1, 0], 1 - Q[:, 1, 1] - Q[:, 2, 2]], 1),
    ], 1)
    pts = [torch.tensor(kw["points"][i], dtype=dd) @ R[i].T + t[i] for i in range(N)]
    ijk = torch.stack(torch.meshgrid(*(torch.arange(D, dtype=dd),) * 3, indexing="ij"), -1).reshape(-1, 3)

    def grid_from(points_t, idx, pitch, origin, trunc):
        idx = torch.tensor(idx.reshape(-1), dtype=torch.long)
        hit = idx >= 0
        f = (points_t[idx.clamp(min=0)] - torch.tensor(origin, dtype=dd)) / float(pitch)
        d = float(pitch) * (f - ijk).norm(dim=1)
        tdf = torch.where(hit, d, torch.full_like(d, float(trunc)))
        return 1 - tdf / float(trunc)

    rew_num = 0; pen_num = 0; pen_den = 0
    for i in range(N):
        sf = r["selfs"][i]
        gs = grid_from(pts[i], sf["indices"], kw["pitch"][i], kw["origin"][i], sf["truncation"])
        surface = gs * torch.tensor(sf["w_surface"].reshape(-1), dtype=dd)
        inside = gs * torch.tensor(sf["w_inside"].reshape(-1), dtype=dd)
        gne = torch.tensor(kw["grid_nontarget_empty"][i].reshape(-1), dtype=dd)
        if r["other_used"][i]:
            o = r["others"][i]
            po = torch.cat([p for j, p in enumerate(pts) if j != i], 0)
            go = grid_from(po, o["indices"], kw["pitch"][i], kw["origin"][i], o["truncation"])
            oin = go * torch.tensor(o["w_inside"].reshape(-1), dtype=dd)
            gne = torch.where(gne >= oin, gne, oin)
        rew_num = rew_num + (surface * torch.tensor(kw["grid_target"][i].reshape(-1), dtype=dd)).sum()
        pen_num = pen_num + (inside * gne).sum()
        pen_den = pen_den + inside.sum()
    loss = pen_num / pen_den - rew_num / float(kw["grid_target"].sum())
    np.testing.assert_allclose(float(loss), r["loss"], rtol=1e-4, atol=1e-5)
    loss.backward()
    np.testing.assert_allclose(r["gt"], t.grad.numpy(), rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(r["gq"], q.grad.numpy(), rtol=2e-3, atol=2e-4)


def test_icc_refine_improves():
    g = golden("icc_forward_contact3")
    kw = _icc_inputs(g)
    q, t, hist = oicc.icc_refine(
        g["transform_init"], n_iter=15, voxel_dim=int(g["voxel_dim"]), sdf_offset=0.02,
        return_history=True, **kw)
    assert np.isfinite(hist).all()
    assert min(hist[5:]) < hist[0]


def test_chainer_adam_first_step():
    # t=1: m=(1-b1)g, v=(1-b2)g^2, alpha_t=alpha*sqrt(1-b2)/(1-b1) -> step = alpha*g/(|g|+eps*...)
    opt = oicc.ChainerAdam((3,), 0.01)
    p = np.zeros(3, F32)
    gvec = np.array([1.0, -2.0, 0.5], F32)
    opt.update(p, gvec)
    np.testing.assert_allclose(p, -0.01 * np.sign(gvec), rtol=1e-4)


def test_average_distance_loss():
    from oracle import loss as oloss
    from oracle import metrics as om
    g = golden("average_distance")
    add = oloss.average_distance(g["points"], g["T1"], g["T2"], symmetric=False)
    adds = oloss.average_distance(g["points"], g["T1"], g["T2"], symmetric=True)
    assert np.array_equal(add, g["ref_add"])                              # reference code, exact
    np.testing.assert_allclose(adds, g["ref_add_s"], rtol=1e-6, atol=1e-7)
    # the reference's own KAT (tests/functions_tests/loss_tests/test_average_distance.py:23-31):
    # loss == metrics.average_distance ADD
    for i in range(g["T2"].shape[0]):
        want = om.average_distance(g["points"].astype(np.float64), g["T1"], g["T2"][i])[0]
        np.testing.assert_allclose(add[i], want, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- training-step oracle
def _train_case(B=2, P=300, seed=0):
    from oracle import cnn as ocnn
    rs = np.random.RandomState(seed)
    w = ocnn.init_weights(21, seed=3)
    values = rs.normal(0, 1, (B, 32, P)).astype(np.float32)
    c = rs.uniform(12, 20, (B, 3, 1))
    d = rs.normal(size=(B, 3, P))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    points = (c + d * rs.uniform(6, 11, (B, 1, P))).astype(np.float32)
    points[:, :, :4] = rs.uniform(-3, 35, (B, 3, 4))              # a few outside the grid
    batch = dict(class_id=(np.arange(B) % 21 + 3).astype(np.int32), values=values, points=points,
                 pitch=np.array([0.0065 + 0.0005 * i for i in range(B)], np.float32),
                 origin=rs.uniform(-0.2, 0.6, (B, 3)).astype(np.float32),
                 grid_nontarget_empty=(rs.uniform(size=(B, 32, 32, 32)) < 0.4))
    qt = rs.normal(size=(B, 4)).astype(np.float32)
    qt /= np.linalg.norm(qt, axis=1, keepdims=True)
    kw = dict(quaternion_true=qt, translation_true=rs.uniform(-0.1, 0.7, (B, 3)).astype(np.float32),
              cad_points=[rs.uniform(-0.05, 0.05, (60, 3)).astype(np.float32) for _ in range(B)],
              symmetric=[bool(i % 2) for i in range(B)])
    return w, batch, kw


def test_training_oracle_forward_is_the_inference_oracle():
    """oracle/cnn_train.py (differentiable torch restatement) == oracle/cnn.py bit for bit on the
    forward pass, so gradients taken through it are gradients of the pinned forward arithmetic."""
    import torch
    from oracle import cnn as ocnn, cnn_train as ct
    w, batch, _ = _train_case(B=2, P=1000, seed=1)
    ref = ocnn.forward(w, n_fg_class=21, bf16=False, **batch)
    with torch.no_grad():
        out = ct.forward(ct.params_from_weights(w, requires_grad=False), n_fg_class=21, **batch)
    for k in ("voxelized", "feat", "rot", "trans", "conf"):
        assert np.array_equal(out[k].numpy(), ref[k]), k


def test_training_oracle_loss_matches_operator_oracles():
    """pose_loss == the reference formula evaluated with the already-pinned NumPy oracles
    (transforms.transformation_matrix, loss.average_distance: model.py:405-441)."""
    import torch
    from oracle import cnn_train as ct, loss as oloss, transforms as otf
    w, batch, kw = _train_case()
    with torch.no_grad():
        out = ct.forward(ct.params_from_weights(w, requires_grad=False), n_fg_class=21, **batch)
        got = float(ct.pose_loss(out, **kw))
    want = 0.0
    B = out["rot"].shape[0]
    for i in range(B):
        q, t, conf = out["rot"][i].numpy(), out["trans"][i].numpy(), out["conf"][i].numpy()
        Tp = otf.transformation_matrix(q, t)
        Tt = otf.transformation_matrix(kw["quaternion_true"][i], kw["translation_true"][i])
        add = oloss.average_distance(kw["cad_points"][i], Tt, Tp, symmetric=kw["symmetric"][i])
        keep = conf > 0
        want += np.mean(add[keep].astype(np.float64) * conf[keep]
                        - ct.LAMBDA_CONFIDENCE * np.log(conf[keep].astype(np.float64)))
    want /= B
    assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (got, want)


def test_training_oracle_gradients_match_finite_differences():
    """float32 autograd gradients == float64 autograd gradients (<=1e-3 of the largest entry) and
    float64 directional derivatives == central finite differences (<=2e-3 relative + 2e-7)."""
    import torch
    from oracle import cnn_train as ct
    w, batch, kw = _train_case()
    l32, g32, _ = ct.loss_and_grads(w, batch, **kw)
    l64, g64, _ = ct.loss_and_grads(w, batch, dtype=torch.float64, **kw)
    assert abs(l32 - l64) <= 1e-5 * max(1.0, abs(l64))
    for k in w:
        scale = np.abs(g64[k]).max()
        assert scale > 0, k                                     # every parameter gets a gradient
        assert np.abs(g32[k] - g64[k]).max() <= 1e-3 * scale, k

    def loss64(wd):
        p = {k: torch.tensor(np.asarray(v, dtype=np.float64)) for k, v in wd.items()}
        with torch.no_grad():
            return float(ct.pose_loss(ct.forward(p, n_fg_class=21, **batch), **kw))

    rs = np.random.RandomState(5)
    for k in ("conv3/W", "conv2_pcd/W", "conv1_occ/W", "conv3_trans/W", "conv4_conf/b"):
        d = rs.normal(size=w[k].shape)
        d /= np.linalg.norm(d)
        eps = 1e-4
        wp, wm = dict(w), dict(w)
        wp[k] = w[k].astype(np.float64) + eps * d
        wm[k] = w[k].astype(np.float64) - eps * d
        fd = (loss64(wp) - loss64(wm)) / (2 * eps)
        an = float((g64[k].astype(np.float64) * d).sum())
        assert abs(fd - an) <= 2e-3 * abs(an) + 2e-7, (k, fd, an)     # 2e-7: fp64 FD noise floor


def test_octree_mapping_oracle_equals_reference_code_over_it():
    """tests/golden/octree_mapping.npz = the reference's own MultiInstanceOctreeMapping code
    (integrate / get_target_grids / get_target_pcds) run over the restated OcTree; the oracle's
    restatement of that class must reproduce it bit for bit (grid assembly, overwrite order)."""
    from oracle import octomap as oc
    g = golden("octree_mapping")
    m = oc.MultiInstanceOctreeMapping()
    for ins, pitch in zip(g["instance_ids"], g["pitches"]):
        m.initialize(int(ins), pitch=float(pitch))
    for n in range(2):
        for ins in g["instance_ids"]:
            m.integrate(int(ins), g[f"label{n}"] == ins, g[f"pcd{n}"], origin=g[f"origin{n}"])
    for ins in g["instance_ids"]:
        cells = m._octrees[int(ins)].cells
        keys = g[f"cells_keys_{ins}"]
        assert len(cells) == len(keys)
        got = np.array([cells[tuple(k)] for k in keys], np.float32)
        assert np.array_equal(got, g[f"cells_logodds_{ins}"])
    for tid, pitch in ((1, 0.006), (3, 0.005)):
        grids = m.get_target_grids(tid, dimensions=(32, 32, 32), pitch=pitch, origin=g[f"grid_origin_{tid}"])
        for name, a in zip(("target", "nontarget", "empty"), grids):
            assert a.dtype == np.float32
            assert np.array_equal(a, g[f"ref_grid_{name}_{tid}"]), (tid, name)
    occ, emp = m.get_target_pcds(2)
    assert np.array_equal(occ, g["ref_pcd_occupied_2"]) and np.array_equal(emp, g["ref_pcd_empty_2"])
