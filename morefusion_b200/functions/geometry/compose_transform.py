"""compose_transform -> mf_compose_transform_fwd (backward is a slice, as in the reference).

API of morefusion/functions/geometry/compose_transform.py:37-48 (class :5-34)."""

import torch

from ... import _lib
from . import _util


class ComposeTransform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, R, t):
        L = _lib.lib()
        _lib.require_cuda(R, t)
        R, t = R.contiguous(), t.contiguous()
        N = R.shape[0]
        T = torch.empty((N, 4, 4), dtype=torch.float32, device=R.device)
        with torch.cuda.device(R.device):
            rc = L.mf_compose_transform_fwd(_lib.ptr(R), _lib.ptr(t), N, _lib.ptr(T), _lib.stream())
        _lib.check(rc, "compose_transform")
        return T

    @staticmethod
    def backward(ctx, gT):
        return gT[:, :3, :3], gT[:, :3, 3]      # compose_transform.py:30-34


def compose_transform(R, t):
    R = _util.as_f32(R)
    t = _util.as_f32(t, R.device)
    squeeze_axis0 = False
    if R.dim() == 2 and t.dim() == 1:
        R = R[None]
        t = t[None]
        squeeze_axis0 = True
    # compose_transform.py:9-16
    _util.expect(R.dim() == 3 and tuple(R.shape[1:3]) == (3, 3), "R.shape == (N, 3, 3)")
    _util.expect(t.dim() == 2 and t.shape[1] == 3, "t.shape == (N, 3)")
    _util.expect(R.shape[0] == t.shape[0], "R.shape[0] == t.shape[0]")
    matrix = ComposeTransform.apply(R, t)
    if squeeze_axis0:
        matrix = matrix[0, :, :]
    return matrix
