"""Differentiable torch-CPU restatement of the 3-D section of singleview_3d.Model and of its
training loss (fp32): the arbiter for the training step (SURVEY.md 8 rows a10-a12 backward).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Forward arithmetic = oracle/cnn.py::forward(bf16=False) (reference model.py:93-164, :239-273),
rewritten with torch ops that autograd can differentiate:
  * _voxelize (model.py:143-164 -> functions/geometry/average_voxelization_3d.py:24-37, :170-206):
    per-voxel sums with index_add_ (CPU: sequential, i.e. ascending point order) divided by the
    counts; gradient = g / count gathered at each point's voxel, none to the points;
  * interpolate_voxel_grid (interpolate_voxel_grid.py:6-59, :132-154): 8-corner gather, corners
    outside the grid skipped, accumulation in corner order; gradient to the grid only
    (interpolate_voxel_grid.py:268 returns None for points).
Loss = Model.loss (model.py:377-441, :475-476) for the "add", "add_s" and "add/add_s" variants:
per object mean(add * conf - lambda * log(conf)) over the points with conf > 0, averaged over
the batch; add = functions.average_distance (functions/loss/average_distance.py:40-85) with
transformation_matrix = translation o quaternion_matrix (quaternion_matrix.py:15-31).  The
occupancy terms of "add+occupancy" (model.py:443-473) are not restated here.
PARITY UNPINNED beyond oracle/cnn.py: the reference has no test that pins Model outputs or
gradients; `tests/test_oracle_golden.py` pins this file's forward to
oracle/cnn.py and its gradients to finite differences.
"""

import numpy as np
import torch
import torch.nn.functional as F

VOXEL_DIM = 32
LAMBDA_CONFIDENCE = 0.015          # Model._lambda_confidence (model.py:15)


def params_from_weights(w, requires_grad=True, dtype=torch.float32):
    """numpy weight dict (oracle.cnn.init_weights) -> dict of torch leaf tensors.  float64 is for
    finite-difference checks of the restatement itself; the arbiter arithmetic is float32."""
    return {k: torch.tensor(np.asarray(v, dtype=np.float32), dtype=dtype,
                            requires_grad=requires_grad)
            for k, v in w.items()}


def _conv1d(p, name, x):
    return F.conv1d(x, p[name + "/W"], p[name + "/b"])


def average_voxelization(values, points, B, D):
    """values [B*P,C] (autograd), points [B*P,3] voxel frame, batch b = n // P.
    origin 0, pitch 1: index = round-half-away((p - 0) / 1), bounds-checked
    (average_voxelization_3d.py:24-37).  Returns [B,C,D,D,D]."""
    N, C = values.shape
    P = N // B
    pts = points.detach()
    tr = torch.trunc(pts)                                      # C roundf: half away from zero
    idx = (tr + torch.sign(pts) * ((pts - tr).abs() >= 0.5).to(pts.dtype)).long()
    ok = ((idx >= 0) & (idx < D)).all(dim=1)
    b = torch.arange(N) // P
    flat = ((b * D + idx[:, 0]) * D + idx[:, 1]) * D + idx[:, 2]
    flat = flat[ok]
    V = B * D * D * D
    sums = torch.zeros(V, C, dtype=values.dtype).index_add_(0, flat, values[ok])
    counts = torch.zeros(V, dtype=values.dtype).index_add_(
        0, flat, torch.ones(flat.numel(), dtype=values.dtype))
    avg = sums / counts.clamp(min=1.0)[:, None]
    return avg.reshape(B, D, D, D, C).permute(0, 4, 1, 2, 3)


def interpolate_voxel_grid(grid, points, B):
    """grid [B,C,X,Y,Z] (autograd), points [B*P,3] in the grid's voxel frame -> [B*P,C]."""
    _, C, X, Y, Z = grid.shape
    N = points.shape[0]
    P = N // B
    pts = points.detach()
    i0 = pts.to(torch.int64)                                   # (int) truncation toward zero
    l = pts - i0.to(pts.dtype)
    h = 1.0 - l
    lx, ly, lz = l[:, 0], l[:, 1], l[:, 2]
    hx, hy, hz = h[:, 0], h[:, 1], h[:, 2]
    ws = [hx * hy * hz, lx * hy * hz, hx * ly * hz, hx * hy * lz,
          lx * ly * hz, hx * ly * lz, lx * hy * lz, lx * ly * lz]
    offs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]
    b = torch.arange(N) // P
    g = grid.permute(0, 2, 3, 4, 1).reshape(B * X * Y * Z, C)
    out = torch.zeros(N, C, dtype=grid.dtype)
    for wj, (dx, dy, dz) in zip(ws, offs):
        ix, iy, iz = i0[:, 0] + dx, i0[:, 1] + dy, i0[:, 2] + dz
        ok = (ix >= 0) & (ix < X) & (iy >= 0) & (iy < Y) & (iz >= 0) & (iz < Z)
        flat = ((b * X + ix.clamp(0, X - 1)) * Y + iy.clamp(0, Y - 1)) * Z + iz.clamp(0, Z - 1)
        out = out + torch.where(ok[:, None], wj[:, None] * g[flat],
                                torch.zeros((), dtype=grid.dtype))
    return out


def forward(p, *, class_id, values, points, pitch, origin, grid_nontarget_empty=None,
            n_fg_class=21):
    """Same signature / outputs as oracle.cnn.forward (rot [B,P,4], trans [B,P,3], conf [B,P])
    but as torch tensors attached to the autograd graph of the parameters `p`."""
    with_occ = "conv1_occ/W" in p
    dt = p["conv3/W"].dtype
    # inputs are float32 data; a float64 run (finite-difference checks) widens them exactly
    values = torch.as_tensor(np.asarray(values, dtype=np.float32)).to(dt)
    points = torch.as_tensor(np.asarray(points, dtype=np.float32)).to(dt)
    B, _, P = values.shape
    D = VOXEL_DIM
    to_center = (D / 2.0 - 0.5) - points
    h_rgb = F.relu(_conv1d(p, "conv1_rgb", values))
    h_pcd = F.relu(_conv1d(p, "conv1_pcd", to_center))
    feat1 = torch.cat((h_rgb, h_pcd), 1)
    h_rgb = F.relu(_conv1d(p, "conv2_rgb", h_rgb))
    h_pcd = F.relu(_conv1d(p, "conv2_pcd", h_pcd))
    feat2 = torch.cat((h_rgb, h_pcd), 1)                                  # [B,144,P]
    pts = points.permute(0, 2, 1).reshape(B * P, 3)
    voxelized = average_voxelization(feat2.permute(0, 2, 1).reshape(B * P, -1), pts, B, D)
    if with_occ:
        g = torch.as_tensor(np.asarray(grid_nontarget_empty).astype(np.float32))[:, None].to(dt)
        h_occ = F.relu(F.conv3d(g, p["conv1_occ/W"], p["conv1_occ/b"], stride=1, padding=1))
        h_occ = F.relu(F.conv3d(h_occ, p["conv2_occ/W"], p["conv2_occ/b"], stride=1, padding=2,
                                dilation=2))
        voxelized = torch.cat([voxelized, h_occ], 1)                      # [B,160,32^3]
    h = F.relu(F.conv3d(voxelized, p["conv3/W"], p["conv3/b"], stride=2, padding=1))
    feat3 = interpolate_voxel_grid(h, pts / 2.0, B).reshape(B, P, 256).permute(0, 2, 1)
    h = F.relu(F.conv3d(h, p["conv4/W"], p["conv4/b"], stride=2, padding=1))
    feat4 = interpolate_voxel_grid(h, pts / 4.0, B).reshape(B, P, 512).permute(0, 2, 1)
    feat = torch.cat((feat1, feat2, feat3, feat4), 1)                     # [B,984,P]
    outs = {}
    for head in ("rot", "trans", "conf"):
        x = feat
        for layer in (1, 2, 3):
            x = F.relu(_conv1d(p, f"conv{layer}_{head}", x))
        outs[head] = _conv1d(p, f"conv4_{head}", x)
    cls_rot = outs["rot"].reshape(B, n_fg_class, 4, P)
    cls_trans = outs["trans"].reshape(B, n_fg_class, 3, P)
    cls_conf = torch.sigmoid(outs["conf"]).reshape(B, n_fg_class, P)
    pitch_t = torch.as_tensor(np.asarray(pitch, dtype=np.float32)).to(dt)
    origin_t = torch.as_tensor(np.asarray(origin, dtype=np.float32)).to(dt)
    pts_cam = points * pitch_t[:, None, None] + origin_t[:, :, None]
    cls_trans = cls_trans * pitch_t[:, None, None, None]
    cls_trans = pts_cam[:, None, :, :] + cls_trans
    fg = torch.as_tensor(np.asarray(class_id)).long() - 1
    ar = torch.arange(B)
    rot = cls_rot[ar, fg]
    rot = rot / (rot.norm(dim=1, keepdim=True) + 1e-5)                     # chainer F.normalize
    return dict(rot=rot.permute(0, 2, 1), trans=cls_trans[ar, fg].permute(0, 2, 1),
                conf=cls_conf[ar, fg], feat=feat, voxelized=voxelized)


def transformation_matrix(q, t):
    """q [N,4] (w,x,y,z), t [N,3] -> [N,4,4] = translation @ rotation
    (quaternion_matrix.py:15-31: q * sqrt(2 / |q|^2), outer product, 9 entries)."""
    n = (q * q).sum(dim=1, keepdim=True)
    qs = q * torch.sqrt(2.0 / n)
    Q = qs[:, :, None] * qs[:, None, :]
    one = torch.ones_like(Q[:, 0, 0])
    zero = torch.zeros_like(one)
    rows = [
        [one - Q[:, 2, 2] - Q[:, 3, 3], Q[:, 1, 2] - Q[:, 3, 0], Q[:, 1, 3] + Q[:, 2, 0], t[:, 0]],
        [Q[:, 1, 2] + Q[:, 3, 0], one - Q[:, 1, 1] - Q[:, 3, 3], Q[:, 2, 3] - Q[:, 1, 0], t[:, 1]],
        [Q[:, 1, 3] - Q[:, 2, 0], Q[:, 2, 3] + Q[:, 1, 0], one - Q[:, 1, 1] - Q[:, 2, 2], t[:, 2]],
        [zero, zero, zero, one],
    ]
    return torch.stack([torch.stack(r, dim=1) for r in rows], dim=1)


def average_distance(cad, T_true, T_pred, symmetric):
    """cad [M,3], T_true [4,4], T_pred [P,4,4] -> [P]  (average_distance.py:40-85).
    ADD-S re-indexes the true points with the nearest neighbour (first minimum) of every
    predicted point; the index is a constant for the gradient."""
    a = cad @ T_true[:3, :3].T + T_true[:3, 3]                             # [M,3]
    b = torch.einsum("pij,mj->pmi", T_pred[:, :3, :3], cad) + T_pred[:, None, :3, 3]
    if symmetric:
        with torch.no_grad():
            d2 = ((b[:, :, None, :] - a[None, None, :, :]) ** 2).sum(-1)   # [P,M,M]
            idx = d2.argmin(dim=2)
        a_sel = a[idx]
    else:
        a_sel = a[None].expand_as(b)
    return torch.sqrt(((a_sel - b) ** 2).sum(-1)).mean(dim=1)


def pose_loss(out, *, quaternion_true, translation_true, cad_points, symmetric,
              lambda_confidence=LAMBDA_CONFIDENCE):
    """out = forward(...); cad_points: list of [M,3] arrays (the 500 sampled CAD points per
    object, model.py:416-418); symmetric: list of bool (model.py:420-428)."""
    B = out["rot"].shape[0]
    dt = out["rot"].dtype
    qt = torch.as_tensor(np.asarray(quaternion_true, dtype=np.float32)).to(dt)
    tt = torch.as_tensor(np.asarray(translation_true, dtype=np.float32)).to(dt)
    loss = 0.0
    for i in range(B):
        T_pred = transformation_matrix(out["rot"][i], out["trans"][i])       # [P,4,4]
        T_true = transformation_matrix(qt[i:i + 1], tt[i:i + 1])[0]
        cad = torch.as_tensor(np.asarray(cad_points[i], dtype=np.float32)).to(dt)
        add = average_distance(cad, T_true, T_pred, bool(symmetric[i]))
        conf = out["conf"][i]
        keep = conf.detach() > 0
        loss = loss + (add[keep] * conf[keep] - lambda_confidence * torch.log(conf[keep])).mean()
    return loss / B


def loss_and_grads(w, batch, *, quaternion_true, translation_true, cad_points, symmetric,
                   n_fg_class=21, dtype=torch.float32):
    """One training evaluation: scalar loss and d loss / d weight for every entry of `w`."""
    p = params_from_weights(w, dtype=dtype)
    out = forward(p, n_fg_class=n_fg_class, **batch)
    loss = pose_loss(out, quaternion_true=quaternion_true, translation_true=translation_true,
                     cad_points=cad_points, symmetric=symmetric)
    loss.backward()
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros_like(w[k]))
             for k, v in p.items()}
    return float(loss.detach()), grads, {k: v.detach().numpy() for k, v in out.items()}
