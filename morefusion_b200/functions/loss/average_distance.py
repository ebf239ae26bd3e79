"""average_distance (ADD / ADD-S pose loss) -> mf_average_distance_{fwd,bwd}.

API of morefusion/functions/loss/average_distance.py:40-85: points [P,3], transform_true [4,4],
transforms_pred [M,4,4], symmetric -> [M].  symmetric=True re-indexes the true points by the
nearest neighbour of every predicted point (geometry.nn, constant w.r.t. the graph, :74-79); the
NN search is matrix-free.  Gradients flow to both transforms (not to the CAD points)."""

import torch

from ... import _lib
from ..geometry import _util


class AverageDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, transform_true, transforms_pred, symmetric):
        L = _lib.lib()
        _lib.require_cuda(points, transform_true, transforms_pred)
        points = points.contiguous().float()
        Tt = transform_true.contiguous().float()
        Tp = transforms_pred.contiguous().float()
        P, M = points.shape[0], Tp.shape[0]
        dev = points.device
        out = torch.empty((M,), dtype=torch.float32, device=dev)
        idx = torch.empty((M, P), dtype=torch.int32, device=dev) if symmetric else None
        with torch.cuda.device(dev):
            rc = L.mf_average_distance_fwd(_lib.ptr(points), P, _lib.ptr(Tt), _lib.ptr(Tp), M,
                                           int(bool(symmetric)), _lib.ptr(out), _lib.ptr(idx),
                                           _lib.stream())
        _lib.check(rc, "average_distance")
        ctx.save_for_backward(points, Tt, Tp, idx if idx is not None else torch.empty(0, device=dev))
        ctx.symmetric = bool(symmetric)
        return out

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        points, Tt, Tp, idx = ctx.saved_tensors
        P, M = points.shape[0], Tp.shape[0]
        dev = points.device
        gout = gout.contiguous().float()
        gTp = torch.empty_like(Tp)
        gTt = torch.empty_like(Tt)
        ws = torch.empty((M, 12), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.mf_average_distance_bwd(
                _lib.ptr(gout), _lib.ptr(points), P, _lib.ptr(Tt), _lib.ptr(Tp), M,
                _lib.ptr(idx) if ctx.symmetric else None, _lib.ptr(gTp), _lib.ptr(gTt),
                _lib.ptr(ws), _lib.stream())
        _lib.check(rc, "average_distance backward")
        return None, gTt, gTp, None


class AverageDistanceBatched(torch.autograd.Function):
    """average_distance over the B objects of a batch in one autograd node / one library call:
    points [B,P,3], transform_true [B,4,4], transforms_pred [B,M,4,4], symmetric: sequence of B
    bools -> [B,M].  Same kernels, same numbers as B calls of AverageDistance."""

    @staticmethod
    def forward(ctx, points, transform_true, transforms_pred, symmetric):
        import ctypes
        L = _lib.lib()
        _lib.require_cuda(points, transform_true, transforms_pred)
        points = points.contiguous().float()
        Tt = transform_true.contiguous().float()
        Tp = transforms_pred.contiguous().float()
        B, P = points.shape[:2]
        M = Tp.shape[1]
        dev = points.device
        sym = (ctypes.c_int32 * B)(*[int(bool(x)) for x in symmetric])
        out = torch.empty((B, M), dtype=torch.float32, device=dev)
        idx = torch.empty((B, M, P), dtype=torch.int32, device=dev) if any(sym) else None
        with torch.cuda.device(dev):
            rc = L.mf_average_distance_fwd_batched(
                _lib.ptr(points), P, _lib.ptr(Tt), _lib.ptr(Tp), M, B, sym, _lib.ptr(out),
                _lib.ptr(idx), _lib.stream())
        _lib.check(rc, "average_distance (batched)")
        ctx.save_for_backward(points, Tt, Tp, idx if idx is not None else torch.empty(0, device=dev))
        ctx.sym = sym
        return out

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        points, Tt, Tp, idx = ctx.saved_tensors
        B, P = points.shape[:2]
        M = Tp.shape[1]
        dev = points.device
        gout = gout.contiguous().float()
        gTp = torch.empty_like(Tp)
        gTt = torch.empty_like(Tt)
        ws = torch.empty((B, M, 12), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.mf_average_distance_bwd_batched(
                _lib.ptr(gout), _lib.ptr(points), P, _lib.ptr(Tt), _lib.ptr(Tp), M, B, ctx.sym,
                _lib.ptr(idx) if idx.numel() else None, _lib.ptr(gTp), _lib.ptr(gTt), _lib.ptr(ws),
                _lib.stream())
        _lib.check(rc, "average_distance backward (batched)")
        return None, gTt, gTp, None


def average_distance_batched(points, transform_true, transforms_pred, symmetric):
    return AverageDistanceBatched.apply(points, transform_true, transforms_pred, tuple(symmetric))


def average_distance(points, transform_true, transforms_pred, symmetric=False):
    points = _util.as_f32(points)
    transform_true = _util.as_f32(transform_true, points.device)
    transforms_pred = _util.as_f32(transforms_pred, points.device)
    n_points = points.shape[0]
    n_pred = transforms_pred.shape[0]
    assert tuple(points.shape) == (n_points, 3)
    assert tuple(transform_true.shape) == (4, 4)
    assert tuple(transforms_pred.shape) == (n_pred, 4, 4)
    return AverageDistance.apply(points, transform_true, transforms_pred, symmetric)
