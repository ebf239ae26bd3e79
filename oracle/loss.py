"""NumPy restatement of functions.average_distance (ADD / ADD-S loss).  TEST INFRASTRUCTURE ONLY.

/root/reference/morefusion/functions/loss/average_distance.py:40-85; nearest neighbour of the
symmetric branch = argmin over the dense fp32 squared-distance matrix of
geometry/knn/cuComputeDistanceGlobal.cu:20-86 + nn.py:48 (first minimum), which agrees with the
CPU path's KD-tree (nn.py:11-14) except on exact distance ties."""

import numpy as np

from . import transforms as tfm

F32 = np.float32


def nn_first_min(ref, query):
    d = (ref[:, None, :].astype(F32) - query[None, :, :].astype(F32))
    d2 = ((d[..., 0] * d[..., 0]).astype(F32) + (d[..., 1] * d[..., 1]).astype(F32)).astype(F32)
    d2 = (d2 + (d[..., 2] * d[..., 2]).astype(F32)).astype(F32)
    return np.argmin(d2, axis=0)


def _tf(points, T):
    # explicit fp32 mul/add order (x*T0 + y*T1) + z*T2 + T3, no FMA
    p = points.astype(F32)
    T = np.asarray(T, dtype=F32)
    out = np.empty(T.shape[:-2] + p.shape, dtype=F32)
    for r in range(3):
        acc = (T[..., r, 0, None] * p[:, 0]).astype(F32)
        acc = (acc + (T[..., r, 1, None] * p[:, 1]).astype(F32)).astype(F32)
        acc = (acc + (T[..., r, 2, None] * p[:, 2]).astype(F32)).astype(F32)
        out[..., r] = (acc + T[..., r, 3, None]).astype(F32)
    return out


def average_distance(points, transform_true, transforms_pred, symmetric=False, return_grads=False,
                     gout=None):
    points = np.asarray(points, dtype=F32)
    Tt = np.asarray(transform_true, dtype=F32)
    Tp = np.asarray(transforms_pred, dtype=F32)
    n_points, n_pred = points.shape[0], Tp.shape[0]
    a = _tf(points, Tt)                  # [P,3]
    b = _tf(points, Tp)                  # [M,P,3]
    if symmetric:
        idx = nn_first_min(a, b.reshape(n_pred * n_points, 3)).reshape(n_pred, n_points)
    else:
        idx = np.tile(np.arange(n_points), (n_pred, 1))
    diff = (a[idx] - b).astype(F32)
    d = np.sqrt((diff ** 2).sum(axis=2, dtype=F32)).astype(F32)
    out = d.mean(axis=1, dtype=np.float64).astype(F32)
    if not return_grads:
        return out
    gout = np.ones(n_pred, F32) if gout is None else np.asarray(gout, F32)
    g_b = (-(diff.astype(np.float64)) / d[..., None].astype(np.float64)) * (gout[:, None, None] / n_points)
    ph = np.concatenate([points.astype(np.float64), np.ones((n_points, 1))], 1)
    gTp = np.zeros((n_pred, 4, 4))
    gTp[:, :3, :] = np.einsum("mpi,pj->mij", g_b, ph)
    gTt = np.zeros((4, 4))
    gTt[:3, :] = np.einsum("mpi,mpj->ij", -g_b, ph[idx])
    return out, gTp.astype(F32), gTt.astype(F32), idx
