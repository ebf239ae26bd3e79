"""max_voxelization_3d -> mf_max_voxelization_3d_{fwd,bwd}.

API of morefusion/functions/geometry/max_voxelization_3d.py:188-210 (class :8-185)."""

import torch

from ... import _lib
from . import _util


class MaxVoxelization3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, points, batch_indices, intensities, batch_size, origin, pitch,
                dimensions):
        L = _lib.lib()
        _lib.require_cuda(values, points, batch_indices, intensities)
        values, points = values.contiguous(), points.contiguous()
        batch_indices, intensities = batch_indices.contiguous(), intensities.contiguous()
        N, C = values.shape
        X, Y, Z = dimensions
        B = int(batch_size)
        dev = values.device
        with torch.cuda.device(dev):
            matrix = torch.empty((B, C, X, Y, Z), dtype=torch.float32, device=dev)
            indices = torch.empty((B, X, Y, Z), dtype=torch.int32, device=dev)
            flags = torch.zeros(1, dtype=torch.int32, device=dev)
            ws = _util.workspace(L.mf_max_voxelization_3d_workspace_bytes(B, X, Y, Z), dev)
            rc = L.mf_max_voxelization_3d_fwd(
                _lib.ptr(values), _lib.ptr(points), _lib.ptr(batch_indices),
                _lib.ptr(intensities), N, C, B, *origin, pitch, X, Y, Z,
                _lib.ptr(matrix), _lib.ptr(indices), _lib.ptr(ws), ws.numel(),
                _lib.ptr(flags), _lib.stream())
        _lib.check(rc, "max_voxelization_3d")
        _util.raise_on_flags(flags)
        ctx.save_for_backward(indices)
        ctx.geom = (N, C, B, dimensions)
        ctx.mark_non_differentiable(indices)
        return matrix, indices

    @staticmethod
    def backward(ctx, gmatrix, _gi):
        L = _lib.lib()
        (indices,) = ctx.saved_tensors
        N, C, B, (X, Y, Z) = ctx.geom
        gmatrix = gmatrix.contiguous()
        gvalues = torch.empty((N, C), dtype=torch.float32, device=gmatrix.device)
        with torch.cuda.device(gmatrix.device):
            rc = L.mf_max_voxelization_3d_bwd(
                _lib.ptr(gmatrix), _lib.ptr(indices), N, C, B, X, Y, Z, _lib.ptr(gvalues),
                _lib.stream())
        _lib.check(rc, "max_voxelization_3d backward")
        return gvalues, None, None, None, None, None, None, None


def max_voxelization_3d(
    values, points, batch_indices, intensities, *, batch_size, origin, pitch, dimensions,
    return_indices=False,
):
    _util.check_dimensions(dimensions)
    values = _util.as_tensor(values)
    points = _util.as_tensor(points, values.device)
    batch_indices = _util.as_tensor(batch_indices, values.device)
    intensities = _util.as_tensor(intensities, values.device)
    _util.check_voxelization_types(values, points, batch_indices)
    voxelized, indices = MaxVoxelization3D.apply(
        values, points, batch_indices, intensities, batch_size, _util.origin3(origin),
        _util.scalar32(pitch), dimensions)
    if return_indices:
        return voxelized, indices
    return voxelized
