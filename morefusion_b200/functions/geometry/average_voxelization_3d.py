"""average_voxelization_3d on torch CUDA tensors -> mf_average_voxelization_3d_{fwd,bwd}.

API of morefusion/functions/geometry/average_voxelization_3d.py:223-244 (class :7-220,
base voxelization_3d.py:5-32): same name, positional order, keyword-only arguments,
output shapes/dtypes, ValueError("points include nan"), ValueError for bad `dimensions`;
gradient flows to `values` only (:145, :220)."""

import torch

from ... import _lib
from . import _util


class AverageVoxelization3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, points, batch_indices, batch_size, origin, pitch, dimensions):
        L = _lib.lib()
        _lib.require_cuda(values, points, batch_indices)
        values = values.contiguous()
        points = points.contiguous()
        batch_indices = batch_indices.contiguous()
        N, C = values.shape
        X, Y, Z = dimensions
        B = int(batch_size)
        dev = values.device
        with torch.cuda.device(dev):
            matrix = torch.empty((B, C, X, Y, Z), dtype=torch.float32, device=dev)
            counts = torch.empty((B, X, Y, Z), dtype=torch.int32, device=dev)
            nws = L.mf_average_voxelization_3d_workspace_bytes(N)
            ws = _util.workspace(nws, dev)
            rc = L.mf_average_voxelization_3d_fwd(
                _lib.ptr(values), _lib.ptr(points), _lib.ptr(batch_indices), N, C, B,
                *origin, pitch, X, Y, Z, _lib.ptr(matrix), _lib.ptr(counts),
                _lib.ptr(ws), ws.numel(), None, _lib.stream())
        _lib.check(rc, "average_voxelization_3d")
        off = L.mf_average_voxelization_3d_flags_offset()
        _util.raise_on_flags(ws[off:off + 4].view(torch.int32))
        ctx.save_for_backward(points, batch_indices, counts)
        ctx.geom = (B, origin, pitch, dimensions)
        ctx.mark_non_differentiable(counts)
        return matrix, counts

    @staticmethod
    def backward(ctx, gmatrix, _gcounts):
        L = _lib.lib()
        points, batch_indices, counts = ctx.saved_tensors
        B, origin, pitch, (X, Y, Z) = ctx.geom
        gmatrix = gmatrix.contiguous()
        N = points.shape[0]
        C = gmatrix.shape[1]
        gvalues = torch.empty((N, C), dtype=torch.float32, device=gmatrix.device)
        with torch.cuda.device(gmatrix.device):
            rc = L.mf_average_voxelization_3d_bwd(
                _lib.ptr(gmatrix), _lib.ptr(counts), _lib.ptr(points), _lib.ptr(batch_indices),
                N, C, B, *origin, pitch, X, Y, Z, _lib.ptr(gvalues), _lib.stream())
        _lib.check(rc, "average_voxelization_3d backward")
        return gvalues, None, None, None, None, None, None


def average_voxelization_3d(
    values, points, batch_indices, *, batch_size, origin, pitch, dimensions,
    return_counts=False,
):
    _util.check_dimensions(dimensions)
    values = _util.as_tensor(values)
    points = _util.as_tensor(points, values.device)
    batch_indices = _util.as_tensor(batch_indices, values.device)
    _util.check_voxelization_types(values, points, batch_indices)
    voxel, counts = AverageVoxelization3D.apply(
        values, points, batch_indices, batch_size, _util.origin3(origin),
        _util.scalar32(pitch), dimensions)
    if return_counts:
        return voxel, counts
    return voxel
