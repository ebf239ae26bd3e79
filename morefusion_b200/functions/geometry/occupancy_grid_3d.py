"""occupancy_grid_3d -> mf_occupancy_grid_3d_{fwd,bwd} (matrix-free).

API of morefusion/functions/geometry/occupancy_grid_3d.py:77-85 (class :7-74)."""

import numpy as np
import torch

from ... import _lib
from . import _util


class OccupancyGrid3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, pitch, origin, dims, threshold):
        L = _lib.lib()
        _lib.require_cuda(points)
        points = points.contiguous()
        X, Y, Z = dims
        dev = points.device
        grid = torch.empty((X, Y, Z), dtype=torch.float32, device=dev)
        dmin = torch.empty((X, Y, Z), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.mf_occupancy_grid_3d_fwd(
                _lib.ptr(points), points.shape[0], pitch, *origin, X, Y, Z, threshold,
                _lib.ptr(grid), _lib.ptr(dmin), _lib.stream())
        _lib.check(rc, "occupancy_grid_3d")
        ctx.save_for_backward(points, dmin)
        ctx.geom = (pitch, origin, dims, threshold)
        return grid

    @staticmethod
    def backward(ctx, ggrid):
        L = _lib.lib()
        points, dmin = ctx.saved_tensors
        pitch, origin, (X, Y, Z), threshold = ctx.geom
        ggrid = ggrid.contiguous()
        gpoints = torch.empty_like(points)
        with torch.cuda.device(points.device):
            rc = L.mf_occupancy_grid_3d_bwd(
                _lib.ptr(ggrid), _lib.ptr(dmin), _lib.ptr(points), points.shape[0], pitch,
                *origin, X, Y, Z, threshold, _lib.ptr(gpoints), _lib.stream())
        _lib.check(rc, "occupancy_grid_3d backward")
        return gpoints, None, None, None, None


def occupancy_grid_3d(points, *, pitch, origin, dims, threshold=1):
    points = _util.as_tensor(points)
    # occupancy_grid_3d.py:8-29
    pitch_a = np.asarray(_util.scalar32(pitch), dtype=np.float32)
    assert pitch_a.ndim == 0
    dims_a = np.asarray(dims)
    assert dims_a.shape == (3,)
    _util.expect(points.dtype == torch.float32, "points.dtype == float32")
    _util.expect(points.dim() == 2 and points.shape[1] == 3, "points.shape == (P, 3)")
    return OccupancyGrid3D.apply(
        points, float(pitch_a), _util.origin3(origin), tuple(int(d) for d in dims_a),
        float(threshold))
