"""ADD / ADD-S of one pose pair per object on the matrix-free nearest-neighbour kernel.

API of morefusion/metrics/average_distance.py:22-35: ``average_distance(points, transform1,
transform2, translate=True)`` with three lists of equal length, returning two float64 arrays
(adds, add_ss).  ADD = mean |T1 x - T2 x|; ADD-S = mean over x of |T1 x - NN_{T2 X}(T1 x)|: the
KD-tree of the reference (:14-16) is built over the SECOND transform's points and queried with
the first's.  The CUDA kernel (csrc/loss.cu, mf_average_distance_fwd) searches, for every point
under its `transforms_pred` argument, the nearest point under `transform_true` -- so the metric
passes transform2 as `true` and transform1 as the single `pred`."""

import numpy as np
import torch

from .. import _lib


def _one(points, T1, T2, translate, dev):
    L = _lib.lib()
    p = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32), device=dev)
    assert p.dim() == 2 and p.shape[1] == 3
    Ts = []
    for T in (T1, T2):
        T = np.array(T.detach().cpu().numpy() if isinstance(T, torch.Tensor) else T,
                     dtype=np.float32)
        assert T.shape == (4, 4)
        if not translate:
            T[:3, 3] = 0
        Ts.append(torch.as_tensor(T, device=dev).contiguous())
    t1, t2 = Ts
    n = p.shape[0]
    out = torch.empty(2, dtype=torch.float32, device=dev)
    nn_idx = torch.empty((1, n), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.mf_average_distance_fwd(_lib.ptr(p), n, _lib.ptr(t2), _lib.ptr(t1.reshape(1, 4, 4)),
                                       1, 0, _lib.ptr(out[0:1]), None, _lib.stream())
        _lib.check(rc, "average_distance (ADD)")
        rc = L.mf_average_distance_fwd(_lib.ptr(p), n, _lib.ptr(t2), _lib.ptr(t1.reshape(1, 4, 4)),
                                       1, 1, _lib.ptr(out[1:2]), _lib.ptr(nn_idx), _lib.stream())
        _lib.check(rc, "average_distance (ADD-S)")
    return out


def average_distance(points, transform1, transform2, translate=True, device=None):
    assert isinstance(points, list)
    batch_size = len(points)
    assert len(transform1) == batch_size
    assert len(transform2) == batch_size
    if not torch.cuda.is_available():
        raise RuntimeError("morefusion_b200.metrics runs on CUDA only (no CPU fallback)")
    dev = torch.device(device or "cuda")
    outs = [_one(points[i], transform1[i], transform2[i], translate, dev)
            for i in range(batch_size)]
    adds = np.zeros((batch_size,), dtype=float)
    add_ss = np.zeros((batch_size,), dtype=float)
    if outs:
        res = torch.stack(outs).cpu().numpy()          # one device->host read for the batch
        adds[:], add_ss[:] = res[:, 0], res[:, 1]
    return adds, add_ss
