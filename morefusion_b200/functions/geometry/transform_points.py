"""transform_points -> mf_transform_points_{fwd,bwd} (one fused kernel instead of
concat-ones + batched GEMM + transpose + slice).

API of morefusion/functions/geometry/transform_points.py:6-30: points [P,3], transform
[M,4,4] -> [M,P,3]; a 2-D transform gives [P,3]."""

import torch

from ... import _lib
from . import _util


class TransformPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, transform):
        L = _lib.lib()
        _lib.require_cuda(points, transform)
        points, transform = points.contiguous(), transform.contiguous()
        P, M = points.shape[0], transform.shape[0]
        out = torch.empty((M, P, 3), dtype=torch.float32, device=points.device)
        with torch.cuda.device(points.device):
            rc = L.mf_transform_points_fwd(_lib.ptr(points), P, _lib.ptr(transform), M,
                                           _lib.ptr(out), _lib.stream())
        _lib.check(rc, "transform_points")
        ctx.save_for_backward(points, transform)
        return out

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        points, transform = ctx.saved_tensors
        gout = gout.contiguous()
        P, M = points.shape[0], transform.shape[0]
        gp = torch.empty_like(points) if ctx.needs_input_grad[0] else None
        gT = torch.empty_like(transform) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(points.device):
            rc = L.mf_transform_points_bwd(_lib.ptr(gout), _lib.ptr(points), P,
                                           _lib.ptr(transform), M, _lib.ptr(gp), _lib.ptr(gT),
                                           _lib.stream())
        _lib.check(rc, "transform_points backward")
        return gp, gT


def transform_points(points, transform):
    points = _util.as_f32(points)
    transform = _util.as_f32(transform, points.device)
    N = points.shape[0]
    assert tuple(points.shape) == (N, 3)
    squeeze_axis0 = False
    if transform.dim() == 2:
        transform = transform[None]
        squeeze_axis0 = True
    M = transform.shape[0]
    assert tuple(transform.shape) == (M, 4, 4)
    out = TransformPoints.apply(points, transform)
    if squeeze_axis0:
        out = out[0, :, :]
    return out
