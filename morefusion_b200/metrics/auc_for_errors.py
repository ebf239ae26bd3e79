"""Generic accuracy-vs-threshold AUC (morefusion/metrics/auc_for_errors.py:5-25): accuracy at
`nstep` thresholds in [0, max_threshold], trapezoidal area scaled to [0, 1]."""

import numpy as np


def auc_for_errors(errors, max_threshold, *, nstep=1000, return_xy=False):
    errors = np.sort(np.asarray(errors, dtype=float))
    assert errors.ndim == 1
    assert errors.min() >= 0, f"min of errors must be >=0: {errors.min()}"
    x = np.linspace(0, max_threshold, nstep)
    y = np.searchsorted(errors, x, side="right") / errors.size     # fraction with error <= x
    auc = float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2.0)) / (1.0 * max_threshold)
    return (auc, x, y) if return_xy else auc
