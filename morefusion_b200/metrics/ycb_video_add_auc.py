"""YCB-Video area under the accuracy-threshold curve (host statistics over a handful of scalars).

API of morefusion/metrics/ycb_video_add_auc.py:5-51 (YCB_Video_toolbox plot_accuracy_keyframe.m):
errors above `max_value` count as misses, accuracy = rank / n, VOC-style monotone envelope,
area normalised by `max_value`."""

import numpy as np


def VOCap(rec, prec, max_value=0.1):
    mrec = np.concatenate([[0.0], np.asarray(rec, dtype=float), [max_value]])
    mpre = np.concatenate([[0.0], np.asarray(prec, dtype=float), [prec[-1]]])
    mpre = np.maximum.accumulate(mpre)                 # monotone envelope
    step = np.flatnonzero(mrec[1:] != mrec[:-1]) + 1   # where the threshold changes value
    return float(np.sum((mrec[step] - mrec[step - 1]) * mpre[step]) / max_value)


def ycb_video_add_auc(adds, *, max_value=0.1, return_xy=False):
    adds = np.asarray(adds, dtype=float)
    assert adds.ndim == 1
    assert adds.min() >= 0, f"min of adds must be >=0: {adds.min()}"
    n = adds.size
    d = np.sort(np.where(adds > max_value, np.inf, adds))
    accuracy = np.arange(1, n + 1, dtype=float) / n
    hit = np.isfinite(d)
    if hit.any():
        d, accuracy = d[hit], accuracy[hit]
        auc = VOCap(d, accuracy, max_value=max_value)
        x = np.r_[0, d, max_value]
        y = np.r_[0, accuracy, accuracy[-1]]
    else:
        auc = 0
        x = np.array([0, max_value], dtype=float)
        y = np.array([0, 0], dtype=float)
    return (auc, x, y) if return_xy else auc
