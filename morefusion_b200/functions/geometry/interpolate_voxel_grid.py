"""interpolate_voxel_grid -> mf_interpolate_voxel_grid_{fwd,bwd}.

API of morefusion/functions/geometry/interpolate_voxel_grid.py:271-272 (class :116-268):
voxelized [B,C,X,Y,Z] f32, points [P,3] f32 in voxel units, batch_indices [P] i32 ->
values [P,C]; gradient to `voxelized` only (:268)."""

import torch

from ... import _lib
from . import _util


class InterpolateVoxelGrid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, voxelized, points, batch_indices, channels_last=False):
        L = _lib.lib()
        _lib.require_cuda(voxelized, points, batch_indices)
        voxelized, points = voxelized.contiguous(), points.contiguous()
        batch_indices = batch_indices.contiguous()
        if channels_last:
            B, X, Y, Z, C = voxelized.shape
        else:
            B, C, X, Y, Z = voxelized.shape
        P = points.shape[0]
        values = torch.empty((P, C), dtype=torch.float32, device=voxelized.device)
        with torch.cuda.device(voxelized.device):
            rc = L.mf_interpolate_voxel_grid_fwd(
                _lib.ptr(voxelized), _lib.ptr(points), _lib.ptr(batch_indices), P, B, C, X, Y, Z,
                int(channels_last), _lib.ptr(values), _lib.stream())
        _lib.check(rc, "interpolate_voxel_grid")
        ctx.save_for_backward(points, batch_indices)
        ctx.shape = (B, C, X, Y, Z, bool(channels_last))
        return values

    @staticmethod
    def backward(ctx, gvalues):
        L = _lib.lib()
        points, batch_indices = ctx.saved_tensors
        B, C, X, Y, Z, cl = ctx.shape
        gvalues = gvalues.contiguous()
        shape = (B, X, Y, Z, C) if cl else (B, C, X, Y, Z)
        gvox = torch.empty(shape, dtype=torch.float32, device=gvalues.device)
        with torch.cuda.device(gvalues.device):
            rc = L.mf_interpolate_voxel_grid_bwd(
                _lib.ptr(gvalues), _lib.ptr(points), _lib.ptr(batch_indices), points.shape[0],
                B, C, X, Y, Z, int(cl), _lib.ptr(gvox), _lib.stream())
        _lib.check(rc, "interpolate_voxel_grid backward")
        return gvox, None, None, None


def interpolate_voxel_grid(voxelized, points, batch_indices):
    voxelized = _util.as_tensor(voxelized)
    points = _util.as_tensor(points, voxelized.device)
    batch_indices = _util.as_tensor(batch_indices, voxelized.device)
    # interpolate_voxel_grid.py:117-130
    _util.expect(voxelized.dtype == torch.float32, "voxelized.dtype == float32")
    _util.expect(voxelized.dim() == 5, "voxelized.ndim == 5")
    _util.expect(points.dtype == torch.float32, "points.dtype == float32")
    _util.expect(points.dim() == 2 and points.shape[1] == 3, "points.shape == (P, 3)")
    _util.expect(batch_indices.dtype == torch.int32, "batch_indices.dtype == int32")
    _util.expect(batch_indices.dim() == 1, "batch_indices.ndim == 1")
    _util.expect(batch_indices.shape[0] == points.shape[0],
                 "batch_indices.shape[0] == points.shape[0]")
    return InterpolateVoxelGrid.apply(voxelized, points, batch_indices, False)
