"""Argument validation shared by the operator shims (mirrors chainer's check_type_forward)."""

import numpy as np
import torch

from ... import InvalidType
from ... import _lib


def expect(cond, msg):
    if not cond:
        raise InvalidType(msg)


def as_f32(x, device=None, name="array"):
    """Accept torch tensors (kept on their device) and numpy / python data (moved to the
    current CUDA device): the reference accepts numpy, cupy and Variables alike."""
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x), device=device or "cuda")


def origin3(origin):
    """origin -> three python floats rounded to float32, as cupy.asarray(origin, float32)."""
    if isinstance(origin, torch.Tensor):
        origin = origin.detach().cpu().numpy()
    o = np.asarray(origin, dtype=np.float32).reshape(-1)
    if o.shape != (3,):
        raise ValueError("origin must have 3 elements")
    return float(o[0]), float(o[1]), float(o[2])


def scalar32(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().item()
    return float(np.float32(x))


def check_dimensions(dimensions):
    # voxelization_3d.py:11-16 (message kept verbatim, including its "4")
    if not (isinstance(dimensions, tuple) and len(dimensions) == 3
            and all(isinstance(d, int) for d in dimensions)):
        raise ValueError("dimensions must be a tuple of 4 integers")


def check_voxelization_types(values, points, batch_indices):
    # voxelization_3d.py:18-32
    expect(values.dtype == torch.float32, "values.dtype == float32")
    expect(values.dim() == 2, "values.ndim == 2")
    expect(points.dtype == torch.float32, "points.dtype == float32")
    expect(points.dim() == 2 and points.shape[1] == 3, "points.shape == (P, 3)")
    expect(points.shape[0] == values.shape[0], "points.shape[0] == values.shape[0]")
    expect(batch_indices.dtype == torch.int32, "batch_indices.dtype == int32")
    expect(batch_indices.dim() == 1, "batch_indices.ndim == 1")
    expect(batch_indices.shape[0] == values.shape[0], "batch_indices.shape[0] == values.shape[0]")


def raise_on_flags(flags_tensor):
    from ... import config
    if not config.check_nan:
        return
    f = int(flags_tensor.item())          # device->host sync, as in the reference (:47-48)
    if f & 1:
        raise ValueError("points include nan")


_WS = {}


def workspace(nbytes, device):
    """Grow-only per-device scratch buffer (caller-owned memory in the C ABI's terms)."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf
