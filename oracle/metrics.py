"""NumPy restatement of the reference's pose metrics.  TEST INFRASTRUCTURE ONLY.

/root/reference/morefusion/metrics/average_distance.py:6-19 (ADD, ADD-S with a KD-tree over the
SECOND transform's points) and metrics/ycb_video_add_auc.py:5-51 (YCB-Video VOC-style AUC)."""

import numpy as np
import sklearn.neighbors


def _tp(points, T):
    return points @ np.asarray(T)[:3, :3].T + np.asarray(T)[:3, 3]


def average_distance(points, transform1, transform2):
    p1, p2 = _tp(points, transform1), _tp(points, transform2)
    add = np.linalg.norm(p1 - p2, axis=1).mean()
    idx = sklearn.neighbors.KDTree(p2).query(p1, return_distance=False)[:, 0]
    add_s = np.linalg.norm(p1 - p2[idx], axis=1).mean()
    return add, add_s


def ycb_video_add_auc(adds, max_value=0.1):
    adds = np.asarray(adds, dtype=np.float64)
    assert adds.ndim == 1 and adds.min() >= 0
    D = adds.copy()
    D[D > max_value] = np.inf
    d = np.sort(D)
    n = len(d)
    acc = np.cumsum(np.ones((1, n))) / n
    keep = np.isfinite(d)
    if not keep.any():
        return 0.0
    d, acc = d[keep], acc[keep]
    mrec = np.r_[0, d, max_value]
    mpre = np.r_[0, acc, acc[-1]]
    for i in range(1, len(mpre)):
        mpre[i] = max(mpre[i], mpre[i - 1])
    i = np.argwhere(mrec[1:] != mrec[:-1]) + 1
    return float(np.sum((mrec[i] - mrec[i - 1]) * mpre[i]) / max_value)
