"""GPU: the reference-facing methods of singleview_3d.Model (model.py:166-481) and the product
metrics, against NumPy restatements of the reference lines they mirror."""

import numpy as np
import pytest
import torch

from oracle import metrics as ometrics

pytestmark = pytest.mark.gpu
F32 = np.float32


def _model(dev):
    from morefusion_b200 import synthetic
    from morefusion_b200.contrib.singleview_3d.models import Model
    torch.manual_seed(0)
    m = Model(n_fg_class=21, with_occupancy=True).to(dev)
    m.load_reference_weights(synthetic.init_weights(21, seed=1))
    return m.eval()


def _frame(B=3, H=64, W=64, seed=0):
    rs = np.random.RandomState(seed)
    rgb = rs.randint(0, 255, (B, H, W, 3)).astype(np.uint8)
    pcd = np.full((B, H, W, 3), np.nan, F32)
    class_id = np.array([2, 9, 15][:B], np.int32)
    for i in range(B):
        # a patch of valid depth; object 1 has fewer than 1000 valid pixels (padding branch)
        n = 40 if i != 1 else 25
        yy, xx = np.meshgrid(np.arange(8, 8 + n), np.arange(10, 10 + n), indexing="ij")
        z = 0.6 + 0.05 * rs.rand(n, n)
        pcd[i, yy, xx, 0] = (xx - W / 2) * z / 600.0
        pcd[i, yy, xx, 1] = (yy - H / 2) * z / 600.0
        pcd[i, yy, xx, 2] = z
        drop = rs.rand(n, n) < 0.1
        pcd[i, yy[drop], xx[drop], :] = np.nan
    return class_id, rgb, pcd


def test_predict_sampling_defaults_and_frames(cuda_device):
    """model.py:178 mask, :195-229 per-object sampling with RandomState(1234) / padding,
    :197-205 default pitch and origin, :236 voxel-frame transform."""
    from morefusion_b200.contrib.singleview_3d.models.model import YCB_VOXEL_PITCH_32
    m = _model(cuda_device)
    class_id, rgb, pcd = _frame()
    B = len(class_id)
    seen = {}

    def capture(**kw):
        seen.update(kw)
        return m.forward_features(**kw)
    m._features = capture
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        h_rgb = m.pspnet_extractor(m.resnet_extractor(
            torch.as_tensor(rgb, device=cuda_device).permute(0, 3, 1, 2).float())).cpu().numpy()
        gne = np.zeros((B, 32, 32, 32), bool)
        rot, trans, conf = m.predict(class_id=class_id, rgb=rgb, pcd=pcd, grid_nontarget_empty=gne)
    assert rot.shape == (B, 1000, 4) and trans.shape == (B, 1000, 3) and conf.shape == (B, 1000)
    for i in range(B):
        mask = ~np.isnan(pcd[i]).any(axis=2)
        iy, ix = np.where(mask)
        n = int(mask.sum())
        rs = np.random.RandomState(1234)
        keep = rs.permutation(n)[:1000] if n >= 1000 else np.r_[np.arange(n), rs.randint(0, n, 1000 - n)]
        pts = pcd[i, iy[keep], ix[keep]].T                                  # [3,P]
        pitch = F32(YCB_VOXEL_PITCH_32[int(class_id[i])])
        valid = pcd[i, iy, ix]                                              # [n,3]
        s = np.sort(valid, axis=0)
        med = s[n // 2] if n % 2 == 1 else (s[n // 2] + s[n // 2 - 1]) / 2   # extra/_cupy.py:47-62
        origin = (med - pitch * F32(15.5)).astype(F32)
        np.testing.assert_allclose(seen["pitch"][i].item(), pitch, rtol=1e-7)
        np.testing.assert_allclose(seen["origin"][i].cpu().numpy(), origin, rtol=0, atol=1e-7)
        want = ((pts - origin[:, None]) / pitch).astype(F32)
        np.testing.assert_allclose(seen["points"][i].cpu().numpy(), want, rtol=0, atol=2e-4)
        # the extractor tail runs at the sampled pixels only (csrc/extractor_tail.cu): same numbers
        # as the dense up3 / conv1 / log_softmax up to fp32 summation order
        np.testing.assert_allclose(seen["values"][i].cpu().numpy(), h_rgb[i][:, iy[keep], ix[keep]],
                                   rtol=0, atol=3e-5)
    q = rot.cpu().numpy()
    assert np.isfinite(q).all() and np.allclose(np.linalg.norm(q, axis=2), 1.0, atol=1e-3)
    with pytest.raises(IndexError):
        m.predict(class_id=np.array([0, 1, 2], np.int32), rgb=rgb, pcd=pcd, grid_nontarget_empty=gne)


def test_extractor_tail_at_sampled_pixels_equals_dense_path(cuda_device):
    """csrc/extractor_tail.cu against the dense torch layers it replaces (pspnet.py:64-82 +
    model.py:222), including image corners / borders (zero padding of the 3x3 conv, clamped
    bilinear taps), non-square images, and the exact gather when the switch is off."""
    m = _model(cuda_device)
    torch.manual_seed(3)
    # the dense comparison path must be fp32 too (cuDNN convs default to TF32: 3e-4 on its own)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        m.pspnet_extractor.up3.prelu.weight.fill_(0.2)
        for B, Hs, Ws, P in ((2, 32, 32, 1000), (1, 24, 40, 77), (3, 128, 128, 1000)):
            up2 = torch.randn(B, 64, Hs, Ws, device=cuda_device)
            H, W = 2 * Hs, 2 * Ws
            pix = torch.randint(0, H * W, (B, P), device=cuda_device)
            pix[:, :8] = torch.tensor([0, W - 1, (H - 1) * W, H * W - 1, W, 2 * W - 1, 1, H * W - 2],
                                      device=cuda_device)
            got = m._extractor_tail(up2, pix)
            dense = torch.log_softmax(m.pspnet_extractor.conv1(m.pspnet_extractor.up3(up2)), dim=1)
            want = dense.flatten(2).gather(2, pix[:, None, :].expand(B, 32, P))
            assert got.shape == want.shape
            err = (got - want).abs().max().item()
            assert err < 5e-5, (B, Hs, Ws, err)
    class_id, rgb, pcd = _frame()
    gne = np.zeros((3, 32, 32, 32), bool)
    with torch.no_grad():
        a = m.predict(class_id=class_id, rgb=rgb, pcd=pcd, grid_nontarget_empty=gne)
        m.fused_extractor_tail = False
        b = m.predict(class_id=class_id, rgb=rgb, pcd=pcd, grid_nontarget_empty=gne)
    for x, y in zip(a, b):                       # bf16 3-D section downstream: same inputs to 3e-5
        assert (x - y).abs().max().item() < 2e-2


def test_metrics_average_distance_vs_reference_definition(cuda_device):
    from morefusion_b200 import metrics, synthetic
    rs = np.random.RandomState(3)
    models = synthetic.SyntheticYCBModels()
    pts, T1, T2 = [], [], []
    for c in (1, 13, 20):
        pts.append(models.get_pcd(c))
        for Ts in (T1, T2):
            T = np.eye(4)
            T[:3, :3] = synthetic._rot(rs)
            T[:3, 3] = rs.uniform(-0.05, 0.05, 3) + [0, 0, 0.6]
            Ts.append(T.astype(F32))
    T2[0] = T1[0].copy()
    T2[0][:3, 3] += 0.003
    adds, add_ss = metrics.average_distance(pts, T1, T2)
    for i in range(3):
        a, s = ometrics.average_distance(pts[i].astype(np.float64), T1[i], T2[i])
        np.testing.assert_allclose(adds[i], a, rtol=2e-5)
        np.testing.assert_allclose(add_ss[i], s, rtol=2e-5)
    assert adds.dtype == np.float64 and adds.shape == (3,)


def test_evaluate_and_loss_api(cuda_device):
    """evaluate (model.py:325-375) reports ADD/ADD-S of given poses; loss (model.py:377-441)
    is differentiable w.r.t. the predicted quaternions / translations / confidences."""
    m = _model(cuda_device)
    B, P = 2, 1000
    rs = np.random.RandomState(0)
    class_id = np.array([4, 13], np.int32)
    q_true = rs.normal(size=(B, 4)).astype(F32)
    q_true /= np.linalg.norm(q_true, axis=1, keepdims=True)
    t_true = (rs.uniform(-0.1, 0.1, (B, 3)) + [0, 0, 0.6]).astype(F32)
    summary = m.evaluate(class_id=class_id, quaternion_true=q_true, translation_true=t_true,
                         quaternion_pred=q_true, translation_pred=t_true + 0.01)
    vals = [v for k, v in summary.items() if k.startswith("add/")]
    np.testing.assert_allclose(vals, [np.sqrt(3) * 0.01] * 2, rtol=1e-4)
    dev = cuda_device
    q = torch.tensor(q_true[:, None, :] + 0.05 * rs.normal(size=(B, P, 4)), dtype=torch.float32,
                     device=dev, requires_grad=True)
    t = torch.tensor(t_true[:, None, :] + 0.01 * rs.normal(size=(B, P, 3)), dtype=torch.float32,
                     device=dev, requires_grad=True)
    c = torch.tensor(rs.uniform(0.2, 0.9, (B, P)), dtype=torch.float32, device=dev, requires_grad=True)
    np.random.seed(0)
    loss = m.loss(class_id, q_true, t_true, q, t, c)
    loss.backward()
    assert np.isfinite(float(loss)) and q.grad.abs().sum() > 0 and t.grad.abs().sum() > 0
    # d loss / d conf = (add - lambda / conf) / (B * P)
    assert c.grad.abs().sum() > 0


def test_occupancy_registration_recovers_translation(cuda_device):
    """contrib/occupancy_registration.py:62-139: align a box's surface points to its occupancy
    grid from a 1.5-voxel offset; the loss must fall and the translation error shrink."""
    from morefusion_b200.contrib import OccupancyRegistration
    from morefusion_b200 import synthetic
    rs = np.random.RandomState(0)
    D, pitch = 16, 0.01
    half = np.array([0.04, 0.03, 0.02])
    pts = synthetic.surface_points("box", half, 600, rs).astype(F32)
    origin = (-(D / 2.0 - 0.5) * pitch,) * 3
    ijk = np.stack(np.meshgrid(*(np.arange(D),) * 3, indexing="ij"), -1) * pitch + np.array(origin)
    d = synthetic.sdf_primitive("box", half, ijk)
    occupied = (np.abs(d) < 0.75 * pitch).astype(F32)
    unoccupied = (d < -1.5 * pitch).astype(F32)
    T0 = np.eye(4, dtype=F32)
    T0[:3, 3] = [0.015, -0.01, 0.005]
    reg = OccupancyRegistration(pts, np.stack([occupied, unoccupied]), pitch=pitch, origin=origin,
                                threshold=2, transform_init=T0, gpu=cuda_device.index or 0, alpha=0.01)
    T = reg.register(iteration=60)
    assert np.linalg.norm(T[:3, 3]) < 0.6 * np.linalg.norm(T0[:3, 3])


def test_frontend_pointcloud_and_bboxes(cuda_device):
    """geometry.pointcloud_from_depth / masks_to_bboxes vs the reference's NumPy definitions
    (pointcloud_from_depth.py:16-25, masks_to_bboxes.py:27-33)."""
    from morefusion_b200 import geometry
    rs = np.random.RandomState(0)
    H, W = 480, 640
    depth = rs.uniform(0.4, 1.2, (H, W)).astype(F32)
    depth[rs.rand(H, W) < 0.1] = np.nan
    fx, fy, cx, cy = 619.44, 619.32, 326.82, 239.52
    for dtype in ("z", "euclidean"):
        got = geometry.pointcloud_from_depth(depth, fx, fy, cx, cy, dtype)
        c, r = np.meshgrid(np.arange(W), np.arange(H), sparse=True)
        z = depth.astype(np.float64)
        pc = np.dstack((z * (c - cx) / fx, z * (r - cy) / fy, z * np.ones_like(c)))
        if dtype == "euclidean":
            pc = pc * (z / np.linalg.norm(pc, axis=2))[:, :, None]
        assert got.shape == (H, W, 3) and np.array_equal(np.isnan(got), np.isnan(pc))
        np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(pc), rtol=2e-6, atol=1e-7)
    masks = np.zeros((3, H, W), bool)
    masks[0, 10:50, 100:300] = True
    masks[2, 479, 0] = True
    bb = geometry.masks_to_bboxes(masks)
    assert bb.dtype == np.float64
    np.testing.assert_array_equal(bb, [[10, 100, 50, 300], [0, 0, 0, 0], [479, 0, 480, 1]])
    np.testing.assert_array_equal(geometry.masks_to_bboxes(masks[0]), [10, 100, 50, 300])
