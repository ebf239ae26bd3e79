"""torch-CPU restatement of the 3D-CNN section of singleview_3d.Model (fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/morefusion/contrib/singleview_3d/models/model.py:
  layer definitions :62-91, _extract :93-141, _voxelize :143-164,
  heads + pose assembly :239-273.
Conv layers are chainer ``L.ConvolutionND`` -> cuDNN (third-party, absent from
/root/reference): cross-correlation with OIDHW weights, identical in definition to
``torch.nn.functional.conv{1,3}d`` used here -- PARITY UNPINNED for the conv numerics
themselves (no reference test pins Model outputs).
``F.normalize`` is chainer's x / (||x|| + 1e-5).

``bf16=True`` rounds operands to bfloat16 at exactly the points where the CUDA path stores
bf16 (GEMM inputs / stored activations), keeping fp32 accumulation: the tight-tolerance
arbiter for the tensor-core path.  ``bf16=False`` is the reference's fp32 arithmetic.
"""

import numpy as np
import torch
import torch.nn.functional as F

from . import voxel_ops as vo

N_POINT = 1000
VOXEL_DIM = 32


def init_weights(n_fg_class=21, seed=0, with_occupancy=True):
    """Seeded LeCun-normal weights keyed by the reference's link names (model.py:62-91); the
    generator lives with the other synthetic-data helpers so that the product bench does not
    need this package."""
    from morefusion_b200.synthetic import init_weights as gen
    return gen(n_fg_class, seed, with_occupancy)


def _r(x, bf16):
    return x.bfloat16().float() if bf16 else x


def _conv1d(w, name, x, bf16=False, round_in=True):
    W = torch.from_numpy(w[name + "/W"])
    b = torch.from_numpy(w[name + "/b"])
    if bf16:
        W = _r(W, True)
        if round_in:
            x = _r(x, True)
    return F.conv1d(x, W, b)


def forward(w, *, class_id, values, points, pitch, origin, grid_nontarget_empty=None,
            n_fg_class=21, bf16=False, threads=None):
    """values [B,32,P] f32 (per-point RGB features), points [B,3,P] f32 in the voxel frame
    ((cam - origin)/pitch, model.py:236), gne [B,32,32,32].
    Returns dict(rot [B,P,4], trans [B,P,3], conf [B,P], feat [B,984,P], ...)."""
    if threads:
        torch.set_num_threads(threads)
    with_occ = "conv1_occ/W" in w
    values = torch.as_tensor(np.asarray(values, dtype=np.float32))
    points = torch.as_tensor(np.asarray(points, dtype=np.float32))
    B, _, P = values.shape
    D = VOXEL_DIM
    # ---- _extract (model.py:93-141)
    to_center = (D / 2.0 - 0.5) - points
    # the per-point stacks run in fp32 in the CUDA path too (weights/inputs unrounded)
    h_rgb = F.relu(_conv1d(w, "conv1_rgb", values))
    h_pcd = F.relu(_conv1d(w, "conv1_pcd", to_center))
    feat1 = torch.cat((h_rgb, h_pcd), 1)
    h_rgb = F.relu(_conv1d(w, "conv2_rgb", h_rgb))
    h_pcd = F.relu(_conv1d(w, "conv2_pcd", h_pcd))
    feat2 = torch.cat((h_rgb, h_pcd), 1)                                  # [B,144,P]
    # _voxelize (model.py:143-164): origin (0,0,0), pitch 1.0
    bi = np.repeat(np.arange(B, dtype=np.int32), P)
    vals = feat2.permute(0, 2, 1).reshape(B * P, -1).numpy()
    pts = points.permute(0, 2, 1).reshape(B * P, 3).numpy()
    vox, _ = vo.average_voxelization_3d_fwd(
        vals, pts, bi, batch_size=B, origin=(0, 0, 0), pitch=1.0, dimensions=(D, D, D))
    voxelized = torch.from_numpy(vox)
    if with_occ:
        g = torch.as_tensor(np.asarray(grid_nontarget_empty).astype(np.float32))[:, None]
        h_occ = F.relu(F.conv3d(g, torch.from_numpy(w["conv1_occ/W"]),
                                torch.from_numpy(w["conv1_occ/b"]), stride=1, padding=1))
        # the CUDA path runs conv2_occ on tensor cores: bf16 operands, fp32 accumulation
        h_occ = F.relu(F.conv3d(_r(h_occ, bf16), _r(torch.from_numpy(w["conv2_occ/W"]), bf16),
                                torch.from_numpy(w["conv2_occ/b"]), stride=1, padding=2, dilation=2))
        voxelized = torch.cat([voxelized, h_occ], 1)                      # [B,160,32^3]
    W3 = _r(torch.from_numpy(w["conv3/W"]), bf16)
    W4 = _r(torch.from_numpy(w["conv4/W"]), bf16)
    h = F.relu(F.conv3d(_r(voxelized, bf16), W3, torch.from_numpy(w["conv3/b"]), stride=2, padding=1))
    assert h.shape == (B, 256, 16, 16, 16)
    h = _r(h, bf16)
    idx = points.permute(0, 2, 1).reshape(B * P, 3).numpy()
    f3 = vo.interpolate_voxel_grid_fwd(h.numpy(), (idx / np.float32(2.0)).astype(np.float32), bi)
    feat3 = torch.from_numpy(f3).reshape(B, P, 256).permute(0, 2, 1)
    h = F.relu(F.conv3d(h, W4, torch.from_numpy(w["conv4/b"]), stride=2, padding=1))
    assert h.shape == (B, 512, 8, 8, 8)
    h = _r(h, bf16)
    f4 = vo.interpolate_voxel_grid_fwd(h.numpy(), (idx / np.float32(4.0)).astype(np.float32), bi)
    feat4 = torch.from_numpy(f4).reshape(B, P, 512).permute(0, 2, 1)
    feat = torch.cat((feat1, feat2, feat3, feat4), 1)                     # [B,984,P]
    # ---- heads (model.py:239-254)
    outs = {}
    for head in ("rot", "trans", "conf"):
        x = feat
        for layer in (1, 2, 3):
            x = F.relu(_conv1d(w, f"conv{layer}_{head}", x, bf16))
        outs[head] = _conv1d(w, f"conv4_{head}", x, bf16)
    cls_rot = outs["rot"].reshape(B, n_fg_class, 4, P)
    cls_trans = outs["trans"].reshape(B, n_fg_class, 3, P)
    cls_conf = torch.sigmoid(outs["conf"]).reshape(B, n_fg_class, P)
    pitch_t = torch.as_tensor(np.asarray(pitch, dtype=np.float32))
    origin_t = torch.as_tensor(np.asarray(origin, dtype=np.float32))
    pts_cam = points * pitch_t[:, None, None] + origin_t[:, :, None]
    cls_trans = cls_trans * pitch_t[:, None, None, None]
    cls_trans = pts_cam[:, None, :, :] + cls_trans
    fg = torch.as_tensor(np.asarray(class_id)).long() - 1
    ar = torch.arange(B)
    rot = cls_rot[ar, fg]
    trans = cls_trans[ar, fg]
    conf = cls_conf[ar, fg]
    rot = rot / (rot.norm(dim=1, keepdim=True) + 1e-5)                     # chainer F.normalize
    return dict(rot=rot.permute(0, 2, 1).numpy(), trans=trans.permute(0, 2, 1).numpy(),
                conf=conf.numpy(), feat=feat.numpy(), voxelized=voxelized.numpy(),
                out_rot=outs["rot"].numpy())
