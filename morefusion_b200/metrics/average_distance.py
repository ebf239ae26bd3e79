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


N_PARTS = 32          # CTAs per pose: the metric has one pose and thousands of points


def _dev_f32(x, dev):
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=dev, dtype=torch.float32)
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32), device=dev)


def _one(points, T1, T2, translate, dev):
    """(ADD, ADD-S) of one object as a CUDA tensor [2]; inputs may be numpy or CUDA tensors (no
    host round trip for the latter)."""
    L = _lib.lib()
    p = _dev_f32(points, dev).contiguous()
    assert p.dim() == 2 and p.shape[1] == 3
    t1, t2 = _dev_f32(T1, dev), _dev_f32(T2, dev)
    assert t1.shape == (4, 4) and t2.shape == (4, 4)
    if not translate:
        t1, t2 = t1.clone(), t2.clone()
        t1[:3, 3] = 0
        t2[:3, 3] = 0
    t1, t2 = t1.contiguous().reshape(1, 4, 4), t2.contiguous()
    n = p.shape[0]
    parts = torch.empty((2, N_PARTS), dtype=torch.float32, device=dev)
    nn_idx = torch.empty((1, n), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.mf_average_distance_fwd_parts(_lib.ptr(p), n, _lib.ptr(t2), _lib.ptr(t1), 1, 0, N_PARTS,
                                             _lib.ptr(parts[0]), None, _lib.stream())
        _lib.check(rc, "average_distance (ADD)")
        rc = L.mf_average_distance_fwd_parts(_lib.ptr(p), n, _lib.ptr(t2), _lib.ptr(t1), 1, 1, N_PARTS,
                                             _lib.ptr(parts[1]), _lib.ptr(nn_idx), _lib.stream())
        _lib.check(rc, "average_distance (ADD-S)")
    return parts.sum(dim=1)


def average_distance_device(points, transform1, transform2, translate=True, device=None):
    """Same as average_distance but asynchronous: returns a CUDA tensor [B, 2] (ADD, ADD-S) on the
    current stream, no device->host read (the training step reads it when it reports)."""
    dev = torch.device(device or "cuda")
    return torch.stack([_one(points[i], transform1[i], transform2[i], translate, dev)
                        for i in range(len(points))])


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
