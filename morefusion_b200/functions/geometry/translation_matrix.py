"""translation_matrix -> mf_compose_transform_fwd with R = identity.

API of morefusion/functions/geometry/translation_matrix.py:29-39 (class :5-27)."""

import torch

from ... import _lib
from . import _util


class TranslationMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        L = _lib.lib()
        _lib.require_cuda(t)
        t = t.contiguous()
        N = t.shape[0]
        T = torch.empty((N, 4, 4), dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            rc = L.mf_compose_transform_fwd(None, _lib.ptr(t), N, _lib.ptr(T), _lib.stream())
        _lib.check(rc, "translation_matrix")
        return T

    @staticmethod
    def backward(ctx, gT):
        return gT[:, :3, 3]                       # translation_matrix.py:24-27


def translation_matrix(translation):
    translation = _util.as_f32(translation)
    squeeze_axis0 = False
    if translation.dim() == 1:
        translation = translation[None]
        squeeze_axis0 = True
    _util.expect(translation.dim() == 2 and translation.shape[1] == 3, "t.shape == (N, 3)")
    matrix = TranslationMatrix.apply(translation)
    if squeeze_axis0:
        matrix = matrix[0, :, :]
    return matrix
