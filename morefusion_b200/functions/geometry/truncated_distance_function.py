"""truncated_distance_function / pseudo_occupancy_voxelization.

API of morefusion/functions/geometry/truncated_distance_function.py:169-178 and :181-213
(class :6-166).  Gradient flows to `points` only; the pseudo-occupancy weights are
constants (:196-213), exactly as in the reference's graph."""

import torch

from ... import _lib
from . import _util


def _dims3(dims):
    X, Y, Z = (int(d) for d in dims)
    return X, Y, Z


class TruncatedDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, pitch, origin, dims, truncation):
        L = _lib.lib()
        _lib.require_cuda(points)
        points = points.contiguous()
        X, Y, Z = dims
        P = points.shape[0]
        dev = points.device
        tdf = torch.empty((X, Y, Z), dtype=torch.float32, device=dev)
        indices = torch.empty((X, Y, Z), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ws = _util.workspace(L.mf_truncated_distance_function_workspace_bytes(X, Y, Z), dev)
            rc = L.mf_truncated_distance_function_fwd(
                _lib.ptr(points), P, pitch, *origin, X, Y, Z, truncation, _lib.ptr(tdf),
                _lib.ptr(indices), _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "truncated_distance_function")
        ctx.save_for_backward(points, indices)
        ctx.geom = (pitch, origin, dims, truncation)
        ctx.mark_non_differentiable(indices)
        return tdf, indices

    @staticmethod
    def backward(ctx, gtdf, _gi):
        L = _lib.lib()
        points, indices = ctx.saved_tensors
        pitch, origin, (X, Y, Z), truncation = ctx.geom
        gtdf = gtdf.contiguous()
        gpoints = torch.empty_like(points)
        with torch.cuda.device(points.device):
            rc = L.mf_truncated_distance_function_bwd(
                _lib.ptr(gtdf), _lib.ptr(points), _lib.ptr(indices), points.shape[0], pitch,
                *origin, X, Y, Z, truncation, _lib.ptr(gpoints), _lib.stream())
        _lib.check(rc, "truncated_distance_function backward")
        return gpoints, None, None, None, None


def truncated_distance_function(points, *, pitch, origin, dims, truncation,
                                return_indices=False):
    points = _util.as_f32(points)
    # truncated_distance_function.py:13-19
    _util.expect(points.dim() == 2 and points.shape[1] == 3, "points.shape == (P, 3)")
    tdf, indices = TruncatedDistanceFunction.apply(
        points, _util.scalar32(pitch), _util.origin3(origin), _dims3(dims),
        _util.scalar32(truncation))
    if return_indices:
        return tdf, indices
    return tdf


class PseudoOccupancyVoxelization(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, sdf, pitch, origin, dims, threshold, sdf_offset):
        L = _lib.lib()
        _lib.require_cuda(points, sdf)
        points, sdf = points.contiguous(), sdf.contiguous()
        X, Y, Z = dims
        dev = points.device
        f = lambda: torch.empty((X, Y, Z), dtype=torch.float32, device=dev)  # noqa: E731
        grid, surface, inside, tdf, w_surf, w_in = f(), f(), f(), f(), f(), f()
        indices = torch.empty((X, Y, Z), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            ws = _util.workspace(L.mf_pseudo_occupancy_voxelization_workspace_bytes(X, Y, Z), dev)
            rc = L.mf_pseudo_occupancy_voxelization_fwd(
                _lib.ptr(points), _lib.ptr(sdf), points.shape[0], pitch, *origin, X, Y, Z,
                threshold, sdf_offset, _lib.ptr(grid), _lib.ptr(surface), _lib.ptr(inside),
                _lib.ptr(tdf), _lib.ptr(indices), _lib.ptr(w_surf), _lib.ptr(w_in),
                _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "pseudo_occupancy_voxelization")
        ctx.save_for_backward(points, indices, w_surf, w_in)
        import numpy as np
        trunc = float(np.float32(threshold) * np.float32(pitch))
        ctx.geom = (pitch, origin, dims, trunc)
        return grid, surface, inside

    @staticmethod
    def backward(ctx, g_grid, g_surface, g_inside):
        L = _lib.lib()
        points, indices, w_surf, w_in = ctx.saved_tensors
        pitch, origin, (X, Y, Z), trunc = ctx.geom
        # grid = 1 - tdf/trunc; surface = grid*w_surf; inside = grid*w_in (weights constant)
        ggrid = torch.zeros((X, Y, Z), dtype=torch.float32, device=points.device)
        if g_grid is not None:
            ggrid = ggrid + g_grid
        if g_surface is not None:
            ggrid = ggrid + g_surface * w_surf
        if g_inside is not None:
            ggrid = ggrid + g_inside * w_in
        gtdf = (-(ggrid) / trunc).contiguous()
        gpoints = torch.empty_like(points)
        with torch.cuda.device(points.device):
            rc = L.mf_truncated_distance_function_bwd(
                _lib.ptr(gtdf), _lib.ptr(points), _lib.ptr(indices), points.shape[0], pitch,
                *origin, X, Y, Z, trunc, _lib.ptr(gpoints), _lib.stream())
        _lib.check(rc, "pseudo_occupancy_voxelization backward")
        return gpoints, None, None, None, None, None, None


def pseudo_occupancy_voxelization(points, sdf, *, pitch, origin, dims, threshold=1,
                                  sdf_offset=0):
    points = _util.as_f32(points)
    sdf = _util.as_f32(sdf, points.device)
    _util.expect(points.dim() == 2 and points.shape[1] == 3, "points.shape == (P, 3)")
    return PseudoOccupancyVoxelization.apply(
        points, sdf, _util.scalar32(pitch), _util.origin3(origin), _dims3(dims),
        float(threshold), float(sdf_offset))
