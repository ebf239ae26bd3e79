"""Host-side geometry helpers used at link construction time (NumPy, not on the hot path).

The reference initialises ICC parameters with trimesh.transformations.quaternion_from_matrix /
translation_from_matrix (contrib/iterative_collision_check_link.py:21-25)."""

import numpy as np


def quaternion_from_matrix(matrix):
    """Rotation part of a 4x4 (or 3x3) -> unit quaternion (w, x, y, z) with w >= 0.

    Branch-on-largest-diagonal (Shepperd) extraction: numerically stable and, up to
    round-off, the same quaternion trimesh's eigen-decomposition returns."""
    M = np.asarray(matrix, dtype=np.float64)
    m = M[:3, :3]
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0.0:
        s = np.sqrt(tr + 1.0) * 2.0
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s,
                      (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2.0
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s,
                      (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2.0
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s,
                      (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2.0
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s,
                      0.25 * s])
    q /= np.linalg.norm(q)
    if q[0] < 0.0:
        q = -q
    return q


def translation_from_matrix(matrix):
    return np.array(matrix, copy=True)[:3, 3]
