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

Canonical arithmetic (``exact=True``, the default).  The reference leaves three things
undefined: the rounding of F.matmul's 4x4 product (cuBLAS), the order of its fp32 atomicAdd
sums, and chainer's einsum in the quaternion backward.  Adam divides every gradient component
by its own sqrt(v), so summation noise on a nearly-cancelling component becomes an O(alpha)
pose difference within a few dozen iterations -- closed-loop parity (1e-4 after 100
iterations) is only testable against ONE definition of those sums.  The oracle therefore
  * transforms points as x = ((R0*px + R1*py) + R2*pz) + t in fp32 (no FMA),
  * forms every per-voxel backward term in fp32 in a fixed expression order,
  * accumulates the 4 loss sums and the 12 per-object (gR | gt) sums in FP64 from exactly
    representable terms and rounds once to fp32 (order-independent to 1e-16), and
  * evaluates the quaternion backward as explicit left-to-right fp32 sums.
``exact=False`` keeps the first restatement (BLAS matmul, fp32 np.add.at sums, einsum); the two
agree to fp32 round-off per call (tests/test_oracle_golden.py).
"""

import numpy as np

from . import transforms as tfm
from . import voxel_ops as vo

F32 = np.float32


def _sum(x):
    return F32(np.sum(x, dtype=np.float64))


def _rot9(q):
    """(w,x,y,z) fp32 -> 9 fp32 entries, quaternion_matrix.py:15-31 with left-to-right sums."""
    q = [F32(v) for v in q]
    n = F32(F32(F32(q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3])
    s = F32(np.sqrt(F32(F32(2.0) / n)))
    w, x, y, z = (F32(v * s) for v in q)
    xx, yy, zz = F32(x * x), F32(y * y), F32(z * z)
    xy, xz, yz = F32(x * y), F32(x * z), F32(y * z)
    xw, yw, zw = F32(x * w), F32(y * w), F32(z * w)
    one = F32(1.0)
    return np.array([one - yy - zz, xy - zw, xz + yw,
                     xy + zw, one - xx - zz, yz - xw,
                     xz - yw, yz + xw, one - xx - yy], dtype=F32)


def _transform_exact(R9, t, p):
    """x = ((R0*px + R1*py) + R2*pz) + t, fp32, one rounding per operation."""
    p = np.asarray(p, dtype=F32)
    px, py, pz = p[:, 0], p[:, 1], p[:, 2]
    out = np.empty_like(p)
    for a in range(3):
        out[:, a] = ((R9[3 * a] * px + R9[3 * a + 1] * py) + R9[3 * a + 2] * pz) + F32(t[a])
    return out


def _quat_grad_exact(q, gR):
    """quaternion_matrix.py:41-49 + the q*sqrt(2/n) scaling, explicit fp32 expression order
    (the one the CUDA kernel uses)."""
    q = [F32(v) for v in q]
    gR = [F32(v) for v in gR]
    n = F32(F32(F32(q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3])
    s = F32(np.sqrt(F32(F32(2.0) / n)))
    qs = [F32(v * s) for v in q]
    g10, g11 = -gR[5] + gR[7], -gR[4] - gR[8]
    g12, g13 = gR[1] + gR[3], gR[2] + gR[6]
    g20, g22, g23 = gR[2] - gR[6], -gR[0] - gR[8], gR[5] + gR[7]
    g30, g33 = -gR[1] + gR[3], -gR[0] - gR[4]
    gqs = [
        (g10 * qs[1] + g20 * qs[2]) + g30 * qs[3],
        (((g10 * qs[0] + g11 * qs[1]) + g12 * qs[2]) + g13 * qs[3]) + g11 * qs[1],
        ((g20 * qs[0] + g22 * qs[2]) + g23 * qs[3]) + (g12 * qs[1] + g22 * qs[2]),
        (g30 * qs[0] + g33 * qs[3]) + ((g13 * qs[1] + g23 * qs[2]) + g33 * qs[3]),
    ]
    gs = ((gqs[0] * q[0] + gqs[1] * q[1]) + gqs[2] * q[2]) + gqs[3] * q[3]
    gn = gs * (-s / (F32(2.0) * n))
    return np.array([gqs[k] * s + (gn * F32(2.0)) * q[k] for k in range(4)], dtype=F32)


def _accumulate_exact(gsum, owner, cad, dg, xcam, ind, *, pitch, origin, trunc, dims):
    """One grid's backward terms (truncated_distance_function.py:119-145 through x = R p + t):
    for every voxel v with a winner p and d loss/d grid[v] = dg != 0:
        u = (f_p - v) / |f_p - v| ;  a = u * (-dg / trunc)            (fp32)
        gR[owner(p)] += a (x) cad_p ;  gt[owner(p)] += a              (fp64)
    owner[p] = object the winning point belongs to, cad[p] = its CAD-frame coordinates."""
    X, Y, Z = dims
    ind = np.asarray(ind).reshape(-1)
    dg = np.asarray(dg, dtype=F32).reshape(-1)
    hit = np.nonzero((ind >= 0) & (dg != 0))[0]
    if hit.size == 0:
        return
    pid = ind[hit]
    f = ((xcam[pid] - origin[None]) / pitch).astype(F32)
    vx = (hit // (Y * Z)).astype(F32)
    vy = ((hit // Z) % Y).astype(F32)
    vz = (hit % Z).astype(F32)
    d = np.stack([f[:, 0] - vx, f[:, 1] - vy, f[:, 2] - vz], 1).astype(F32)
    n = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]).astype(F32)
    keep = n > 0
    with np.errstate(invalid="ignore", divide="ignore"):
        gtdf = ((-dg[hit]) / trunc).astype(F32)
        a = ((d / n[:, None]).astype(F32) * gtdf[:, None]).astype(F32)
    a, pid = a[keep].astype(np.float64), pid[keep]
    cp = cad[pid].astype(np.float64)
    terms = np.concatenate([a[:, 0:1] * cp, a[:, 0:1], a[:, 1:2] * cp, a[:, 1:2],
                            a[:, 2:3] * cp, a[:, 2:3]], axis=1)        # [n, 12] exact products
    np.add.at(gsum, owner[pid], terms)


def icc_forward_backward(
    quaternion, translation, points, sdf, pitch, origin, grid_target,
    grid_nontarget_empty, *, voxel_dim=32, voxel_threshold=2, sdf_offset=0,
    need_grad=True, exact=True,
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
    if exact:
        pts = [_transform_exact(_rot9(q[i]), t[i], p) for i, p in enumerate(points)]
    else:
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
    if exact:
        cad = [np.asarray(p, dtype=F32) for p in points]
        gsum = np.zeros((N, 12), dtype=np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            for i in range(N):
                s = selfs[i]
                ws, wi = s["w_surface"], s["w_inside"]
                dg = (((wi * gne[i]) * c_in0 - wi * c_in1) - (ws * gt_grid[i]) * c_rw).astype(F32)
                _accumulate_exact(gsum, np.full(sizes[i], i), cad[i], dg, pts[i], s["indices"],
                                  pitch=pitch[i], origin=origin[i], trunc=s["truncation"], dims=dims)
                if other_used[i]:
                    o = others[i]
                    js = [j for j in range(N) if j != i]
                    b1 = np.where(other_wins[i], (o["w_inside"] * inside[i]).astype(F32), F32(0))
                    dgo = (b1 * c_in0).astype(F32)
                    _accumulate_exact(
                        gsum, np.concatenate([np.full(sizes[j], j) for j in js]),
                        np.concatenate([cad[j] for j in js], 0), dgo,
                        np.concatenate([pts[j] for j in js], 0), o["indices"],
                        pitch=pitch[i], origin=origin[i], trunc=o["truncation"], dims=dims)
        g32 = gsum.astype(F32)
        gq = np.stack([_quat_grad_exact(q[i], g32[i, [0, 1, 2, 4, 5, 6, 8, 9, 10]])
                       for i in range(N)])
        out.update(gq=gq, gt=g32[:, [3, 7, 11]].copy(), gsum=gsum)
        return out
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
    voxel_dim=32, voxel_threshold=2, sdf_offset=0, return_history=False, exact=True,
    q0=None, t0=None,
):
    """check_iterative_collision_check_link.py:44-79: link init from 4x4s,
    Adam(alpha), translation alpha *= 0.1, n_iter x (forward, backward, update)."""
    q = np.stack([tfm.quaternion_from_matrix(T) for T in transform_init]).astype(F32)
    t = np.stack([np.asarray(T)[:3, 3] for T in transform_init]).astype(F32)
    # explicit initial parameters: quaternion_from_matrix goes through LAPACK eigh, whose last
    # bits differ between CPUs; the loop amplifies a 1-ulp difference to O(alpha) in 100 steps
    if q0 is not None:
        q = np.array(q0, dtype=F32)
    if t0 is not None:
        t = np.array(t0, dtype=F32)
    oq = ChainerAdam(q.shape, alpha)
    ot = ChainerAdam(t.shape, alpha * translation_alpha_scale)
    hist = []
    for _ in range(n_iter):
        r = icc_forward_backward(
            q, t, points, sdf, pitch, origin, grid_target, grid_nontarget_empty,
            voxel_dim=voxel_dim, voxel_threshold=voxel_threshold, sdf_offset=sdf_offset,
            exact=exact)
        hist.append(float(r["loss"]))
        oq.update(q, r["gq"])
        ot.update(t, r["gt"])
    if return_history:
        return q, t, hist
    return q, t
