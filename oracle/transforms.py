"""NumPy restatement of the reference's rigid-transform operators (fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference files (/root/reference/morefusion/functions/geometry):
  quaternion_matrix.py:15-31 (forward table), :34-51 (backward table),
  :65-78 (normalise*sqrt2, outer); compose_transform.py:18-34;
  translation_matrix.py:13-27; transformation_matrix.py:5-18;
  transform_points.py:6-30.
Quaternions are (w, x, y, z); transforms are row-major 4x4.
"""

import numpy as np

F32 = np.float32


def quaternion_matrix_fwd(q):
    """q [N,4] -> R [N,4,4]. Returns (R, aux) with aux for backward."""
    q = np.asarray(q, dtype=F32)
    squeeze = q.ndim == 1
    if squeeze:
        q = q[None]
    n = (q * q).astype(F32).sum(axis=1, keepdims=True, dtype=F32)
    s = np.sqrt(F32(2.0) / n).astype(F32)
    qs = (q * s).astype(F32)
    Q = (qs[:, :, None] * qs[:, None, :]).astype(F32)       # outer
    N = q.shape[0]
    R = np.tile(np.eye(4, dtype=F32)[None], (N, 1, 1))
    R[:, 0, 0] = 1 - Q[:, 2, 2] - Q[:, 3, 3]
    R[:, 0, 1] = Q[:, 1, 2] - Q[:, 3, 0]
    R[:, 0, 2] = Q[:, 1, 3] + Q[:, 2, 0]
    R[:, 1, 0] = Q[:, 1, 2] + Q[:, 3, 0]
    R[:, 1, 1] = 1 - Q[:, 1, 1] - Q[:, 3, 3]
    R[:, 1, 2] = Q[:, 2, 3] - Q[:, 1, 0]
    R[:, 2, 0] = Q[:, 1, 3] - Q[:, 2, 0]
    R[:, 2, 1] = Q[:, 2, 3] + Q[:, 1, 0]
    R[:, 2, 2] = 1 - Q[:, 1, 1] - Q[:, 2, 2]
    aux = dict(q=q, n=n, s=s, qs=qs)
    if squeeze:
        return R[0], aux
    return R, aux


def quaternion_matrix_bwd(gR, aux):
    """gR [N,4,4] -> gq [N,4], through table(:41-49) -> outer -> scale."""
    gR = np.asarray(gR, dtype=F32)
    if gR.ndim == 2:
        gR = gR[None]
    N = gR.shape[0]
    gQ = np.zeros((N, 4, 4), dtype=F32)
    gQ[:, 1, 0] = -gR[:, 1, 2] + gR[:, 2, 1]
    gQ[:, 1, 1] = -gR[:, 1, 1] - gR[:, 2, 2]
    gQ[:, 1, 2] = gR[:, 0, 1] + gR[:, 1, 0]
    gQ[:, 1, 3] = gR[:, 0, 2] + gR[:, 2, 0]
    gQ[:, 2, 0] = gR[:, 0, 2] - gR[:, 2, 0]
    gQ[:, 2, 2] = -gR[:, 0, 0] - gR[:, 2, 2]
    gQ[:, 2, 3] = gR[:, 1, 2] + gR[:, 2, 1]
    gQ[:, 3, 0] = -gR[:, 0, 1] + gR[:, 1, 0]
    gQ[:, 3, 3] = -gR[:, 0, 0] - gR[:, 1, 1]
    q, n, s, qs = aux["q"], aux["n"], aux["s"], aux["qs"]
    # Q = qs qs^T  ->  gqs = gQ qs + gQ^T qs
    gqs = (np.einsum("nij,nj->ni", gQ, qs) + np.einsum("nji,nj->ni", gQ, qs)).astype(F32)
    # qs = q * s,  s = sqrt(2/n),  n = sum q^2
    gq = (gqs * s).astype(F32)
    gs = (gqs * q).sum(axis=1, keepdims=True, dtype=F32)
    # ds/dn = -0.5 * sqrt(2) * n^-1.5 = -s/(2n)
    gn = (gs * (-s / (F32(2.0) * n))).astype(F32)
    gq = (gq + gn * F32(2.0) * q).astype(F32)
    return gq


def compose_transform(R, t):
    R = np.asarray(R, dtype=F32)
    t = np.asarray(t, dtype=F32)
    squeeze = R.ndim == 2 and t.ndim == 1
    if squeeze:
        R, t = R[None], t[None]
    N = R.shape[0]
    T = np.tile(np.eye(4, dtype=F32)[None], (N, 1, 1))
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    return T[0] if squeeze else T


def translation_matrix(t):
    t = np.asarray(t, dtype=F32)
    squeeze = t.ndim == 1
    if squeeze:
        t = t[None]
    N = t.shape[0]
    T = np.tile(np.eye(4, dtype=F32)[None], (N, 1, 1))
    T[:, :3, 3] = t
    return T[0] if squeeze else T


def transformation_matrix(q, t):
    q = np.asarray(q, dtype=F32)
    t = np.asarray(t, dtype=F32)
    if q.ndim == 2:
        assert q.shape == (q.shape[0], 4) and t.shape == (q.shape[0], 3)
        R, _ = quaternion_matrix_fwd(q)
        return compose_transform(R[:, :3, :3], t)
    assert q.shape == (4,) and t.shape == (3,)
    R, _ = quaternion_matrix_fwd(q[None])
    return compose_transform(R[:, :3, :3], t[None])[0]


def transform_points(points, T):
    """points [P,3], T [M,4,4] | [4,4] -> [M,P,3] | [P,3]."""
    points = np.asarray(points, dtype=F32)
    T = np.asarray(T, dtype=F32)
    P = points.shape[0]
    assert points.shape == (P, 3)
    squeeze = T.ndim == 2
    if squeeze:
        T = T[None]
    assert T.shape == (T.shape[0], 4, 4)
    ph = np.concatenate([points, np.ones((P, 1), dtype=F32)], axis=1)
    out = np.matmul(T, ph.T).transpose(0, 2, 1)[:, :, :3]
    out = np.ascontiguousarray(out, dtype=F32)
    return out[0] if squeeze else out


def quaternion_from_matrix(M):
    """trimesh.transformations.quaternion_from_matrix (third-party, trimesh>=3.5,
    Gohlke's algorithm, isprecise=False): eigenvector of the symmetric 4x4 K
    for the largest eigenvalue, ordered (w,x,y,z), w>=0.  PARITY UNPINNED
    (trimesh absent); any unit quaternion of R is equivalent up to round-off.
    Call site: contrib/iterative_collision_check_link.py:22."""
    M = np.asarray(M, dtype=np.float64)[:4, :4]
    m00, m01, m02 = M[0, 0], M[0, 1], M[0, 2]
    m10, m11, m12 = M[1, 0], M[1, 1], M[1, 2]
    m20, m21, m22 = M[2, 0], M[2, 1], M[2, 2]
    K = np.array([
        [m00 - m11 - m22, 0.0, 0.0, 0.0],
        [m01 + m10, m11 - m00 - m22, 0.0, 0.0],
        [m02 + m20, m12 + m21, m22 - m00 - m11, 0.0],
        [m21 - m12, m02 - m20, m10 - m01, m00 + m11 + m22],
    ])
    K /= 3.0
    w, V = np.linalg.eigh(K)
    q = V[[3, 0, 1, 2], np.argmax(w)]
    if q[0] < 0.0:
        np.negative(q, q)
    return q
