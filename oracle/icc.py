"""NumPy restatement of IterativeCollisionCheckLink + its optimiser loop.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/morefusion/contrib/iterative_collision_check_link.py:31-99
(forward) with the hand-derived reverse pass of the chainer graph it builds
(F.maximum sends the gradient to its FIRST argument on ties; weights carry no
gradient -- truncated_distance_function.py:196-213), and the driver loop of
examples/ycb_video/pose_refinement/check_iterative_collision_check_link.py:44-79.

Adam: ``chainer.optimizers.Adam`` is third-party (requirements.txt:1, unpinned,
absent from /root/reference) -- PARITY UNPINNED.  Restated from chainer v7's
published update rule:  m += (1-b1)(g-m); v += (1-b2)(g*g-v);
alpha_t = alpha*sqrt(1-b2^t)/(1-b1^t);  p -= eta*(alpha_t*m/(sqrt(v)+eps)),
t starting at 1, eps=1e-8, eta=1, b1=0.9, b2=0.999, all in float32.

The forward *loss value* is pinned against the reference's own code run under
oracle/ref_harness (tests/golden/icc_forward_*.npz).
"""

import numpy as np

from . import transforms as tfm
from . import voxel_ops as vo

F32 = np.float32


def _sum(x):
    return F32(np.sum(x, dtype=np.float64))


def icc_forward_backward(
    quaternion, translation, points, sdf, pitch, origin, grid_target,
    grid_nontarget_empty, *, voxel_dim=32, voxel_threshold=2, sdf_offset=0,
    need_grad=True,
):
    """Returns dict(loss, gq [N,4], gt [N,3], + intermediates)."""
    q = np.asarray(quaternion, dtype=F32)
    t = np.asarray(translation, dtype=F32)
    N = q.shape[0]
    dims = (voxel_dim,) * 3
    pitch = np.asarray(pitch, dtype=F32)
    origin = np.asarray(origin, dtype=F32)
    gt_grid = np.asarray(grid_target, dtype=F32)
    gne_in = np.asarray(grid_nontarget_empty, dtype=F32)

    R, qaux = tfm.quaternion_matrix_fwd(q)
    T = tfm.compose_transform(R[:, :3, :3], t)
    pts = [tfm.transform_points(np.asarray(p, dtype=F32), T[i])
           for i, p in enumerate(points)]
    sdf = [np.asarray(s, dtype=F32) for s in sdf]
    sizes = [p.shape[0] for p in pts]

    selfs, others, gne, other_used, other_wins = [], [], [], [], []
    for i in range(N):
        s = vo.pseudo_occupancy_voxelization_fwd(
            pts[i], sdf[i], pitch=pitch[i], origin=origin[i], dims=dims,
            threshold=voxel_threshold, sdf_offset=sdf_offset)
        selfs.append(s)
        g = gne_in[i]
        used, o, wins = False, None, None
        if N > 1:
            po = np.concatenate([p for j, p in enumerate(pts) if j != i], 0)
            so = np.concatenate([p for j, p in enumerate(sdf) if j != i], 0)
            o = vo.pseudo_occupancy_voxelization_fwd(
                po, so, pitch=pitch[i], origin=origin[i], dims=dims,
                threshold=voxel_threshold)
            if not np.isnan(o["inside"]).any():          # :82
                used = True
                wins = ~(g >= o["inside"])               # F.maximum: first arg on ties
                g = np.maximum(g, o["inside"])
        others.append(o)
        other_used.append(used)
        other_wins.append(wins)
        gne.append(g)

    surface = np.stack([s["surface"] for s in selfs])
    inside = np.stack([s["inside"] for s in selfs])
    gne = np.stack(gne)
    rew_num, rew_den = _sum(surface * gt_grid), _sum(gt_grid)
    pen_num, pen_den = _sum(inside * gne), _sum(inside)
    with np.errstate(invalid="ignore", divide="ignore"):
        reward = F32(rew_num / rew_den)
        penalty = F32(pen_num / pen_den)
    loss = F32(penalty - reward)
    out = dict(loss=loss, reward=reward, penalty=penalty, T=T, pts=pts,
               selfs=selfs, others=others, gne=gne, other_used=other_used,
               sums=(rew_num, rew_den, pen_num, pen_den))
    if not need_grad:
        return out

    gpts = [np.zeros_like(p) for p in pts]
    with np.errstate(invalid="ignore", divide="ignore"):
        c_in0 = F32(F32(1) / pen_den)
        c_in1 = F32(pen_num / (pen_den * pen_den))
        c_rw = F32(F32(1) / rew_den)
    for i in range(N):
        s = selfs[i]
        d_surface = -(gt_grid[i] * c_rw)
        d_inside = gne[i] * c_in0 - c_in1
        d_grid = s["w_surface"] * d_surface + s["w_inside"] * d_inside
        d_tdf = (-(d_grid) / s["truncation"]).astype(F32)
        gpts[i] += vo.truncated_distance_function_bwd(
            d_tdf, pts[i], s["indices"], pitch=pitch[i], origin=origin[i], dims=dims)
        if other_used[i]:
            o = others[i]
            d_oin = np.where(other_wins[i], inside[i] * c_in0, F32(0)).astype(F32)
            d_ogrid = o["w_inside"] * d_oin
            d_otdf = (-(d_ogrid) / o["truncation"]).astype(F32)
            po = np.concatenate([p for j, p in enumerate(pts) if j != i], 0)
            gpo = vo.truncated_distance_function_bwd(
                d_otdf, po, o["indices"], pitch=pitch[i], origin=origin[i], dims=dims)
            off = 0
            for j in range(N):
                if j == i:
                    continue
                gpts[j] += gpo[off:off + sizes[j]]
                off += sizes[j]
    gT = np.zeros((N, 4, 4), dtype=F32)
    for i in range(N):
        p = np.asarray(points[i], dtype=np.float64)
        g = gpts[i].astype(np.float64)
        gT[i, :3, :3] = (g.T @ p).astype(F32)
        gT[i, :3, 3] = g.sum(0).astype(F32)
    gR = np.zeros((N, 4, 4), dtype=F32)
    gR[:, :3, :3] = gT[:, :3, :3]
    gq = tfm.quaternion_matrix_bwd(gR, qaux)
    out.update(gq=gq, gt=gT[:, :3, 3].copy(), gpts=gpts, gT=gT)
    return out


class ChainerAdam:
    """chainer.optimizers.Adam for one parameter array (see module header)."""

    def __init__(self, shape, alpha, beta1=0.9, beta2=0.999, eps=1e-8, eta=1.0):
        self.alpha, self.beta1, self.beta2, self.eps, self.eta = alpha, beta1, beta2, eps, eta
        self.m = np.zeros(shape, dtype=F32)
        self.v = np.zeros(shape, dtype=F32)
        self.t = 0

    def alpha_t(self):
        import math
        fix1 = 1.0 - math.pow(self.beta1, self.t)
        fix2 = 1.0 - math.pow(self.beta2, self.t)
        return self.alpha * math.sqrt(fix2) / fix1

    def update(self, param, grad):
        self.t += 1
        g = np.asarray(grad, dtype=F32)
        self.m += F32(1 - self.beta1) * (g - self.m)
        self.v += F32(1 - self.beta2) * (g * g - self.v)
        step = F32(self.alpha_t()) * self.m / (np.sqrt(self.v) + F32(self.eps))
        param -= F32(self.eta) * step.astype(F32)
        return param


def icc_refine(
    transform_init, points, sdf, pitch, origin, grid_target,
    grid_nontarget_empty, *, n_iter=100, alpha=0.01, translation_alpha_scale=0.1,
    voxel_dim=32, voxel_threshold=2, sdf_offset=0, return_history=False,
):
    """check_iterative_collision_check_link.py:44-79: link init from 4x4s,
    Adam(alpha), translation alpha *= 0.1, n_iter x (forward, backward, update)."""
    q = np.stack([tfm.quaternion_from_matrix(T) for T in transform_init]).astype(F32)
    t = np.stack([np.asarray(T)[:3, 3] for T in transform_init]).astype(F32)
    oq = ChainerAdam(q.shape, alpha)
    ot = ChainerAdam(t.shape, alpha * translation_alpha_scale)
    hist = []
    for _ in range(n_iter):
        r = icc_forward_backward(
            q, t, points, sdf, pitch, origin, grid_target, grid_nontarget_empty,
            voxel_dim=voxel_dim, voxel_threshold=voxel_threshold, sdf_offset=sdf_offset)
        hist.append(float(r["loss"]))
        oq.update(q, r["gq"])
        ot.update(t, r["gt"])
    if return_history:
        return q, t, hist
    return q, t
