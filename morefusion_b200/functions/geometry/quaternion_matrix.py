"""quaternion_matrix -> mf_quaternion_matrix_{fwd,bwd}.

API of morefusion/functions/geometry/quaternion_matrix.py:65-78: q (w,x,y,z) [N,4] or [4]
-> homogeneous rotation [N,4,4] or [4,4]; q is normalised inside (q*sqrt(2/|q|^2))."""

import torch

from ... import _lib
from . import _util


class QuaternionMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q):
        L = _lib.lib()
        _lib.require_cuda(q)
        q = q.contiguous()
        N = q.shape[0]
        R = torch.empty((N, 4, 4), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            rc = L.mf_quaternion_matrix_fwd(_lib.ptr(q), N, _lib.ptr(R), _lib.stream())
        _lib.check(rc, "quaternion_matrix")
        ctx.save_for_backward(q)
        return R

    @staticmethod
    def backward(ctx, gR):
        L = _lib.lib()
        (q,) = ctx.saved_tensors
        gR = gR.contiguous()
        gq = torch.empty_like(q)
        with torch.cuda.device(q.device):
            rc = L.mf_quaternion_matrix_bwd(_lib.ptr(gR), _lib.ptr(q), q.shape[0], _lib.ptr(gq),
                                            _lib.stream())
        _lib.check(rc, "quaternion_matrix backward")
        return gq


def quaternion_matrix(quaternion):
    quaternion = _util.as_tensor(quaternion)
    squeeze_axis0 = False
    if quaternion.dim() == 1:
        squeeze_axis0 = True
        quaternion = quaternion[None]
    _util.expect(quaternion.dim() == 2 and quaternion.shape[1] == 4, "quaternion.shape == (N, 4)")
    _util.expect(quaternion.dtype == torch.float32, "quaternion.dtype == float32")
    matrix = QuaternionMatrix.apply(quaternion)
    if squeeze_axis0:
        matrix = matrix[0, :, :]
    return matrix
