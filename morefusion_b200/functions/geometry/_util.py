"""Argument validation shared by the operator shims (mirrors chainer's check_type_forward)."""

import numpy as np
import torch

from ... import InvalidType
from ... import _lib


def expect(cond, msg):
    if not cond:
        raise InvalidType(msg)


def as_tensor(x, device=None):
    """torch tensors are kept as they are; numpy / python data moves to the CUDA device with
    its own dtype.  For the operators whose reference has a ``check_type_forward`` (the dtype is
    then validated, not converted: voxelization_3d.py:18-32, interpolate_voxel_grid.py:117-130)."""
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x), device=device or "cuda")


def as_f32(x, device=None, name="array"):
    """float32 CUDA view of `x` for the operators the reference runs in whatever float dtype it
    is given (no type check there: compose_transform, translation_matrix, transform_points,
    truncated_distance_function, pseudo_occupancy_voxelization, average_distance): float64 /
    float16 tensors and numpy / python data are converted -- the kernels read float32."""
    if isinstance(x, torch.Tensor):
        if x.dtype == torch.float32:
            return x
        if not x.is_floating_point():
            raise InvalidType(f"{name}.dtype.kind == 'f'")
        return x.to(torch.float32)
    a = np.asarray(x)
    if a.dtype.kind not in "fiub":
        raise InvalidType(f"{name}.dtype.kind == 'f'")
    return torch.as_tensor(a.astype(np.float32), device=device or "cuda")


# Scalars / origins given as CUDA tensors have to reach the host once (the C ABI takes them by
# value).  The value is cached per (storage, version): calling an operator in a loop with the
# same pitch / origin tensor -- the ROS node pattern -- costs one device->host read in total, not
# one per call; an in-place update of the tensor bumps its version and refreshes the entry.
_HOST_CACHE = {}


def _host_values(t):
    if not t.is_cuda:
        return t.detach().numpy()
    key = (t.data_ptr(), t._version, tuple(t.shape), t.dtype)
    v = _HOST_CACHE.get(key)
    if v is None:
        if len(_HOST_CACHE) > 256:
            _HOST_CACHE.clear()
        v = t.detach().cpu().numpy()
        _HOST_CACHE[key] = v
    return v


def origin3(origin):
    """origin -> three python floats rounded to float32, as cupy.asarray(origin, float32)."""
    if isinstance(origin, torch.Tensor):
        origin = _host_values(origin)
    o = np.asarray(origin, dtype=np.float32).reshape(-1)
    if o.shape != (3,):
        raise ValueError("origin must have 3 elements")
    return float(o[0]), float(o[1]), float(o[2])


def scalar32(x):
    if isinstance(x, torch.Tensor):
        x = _host_values(x).reshape(-1)[0]
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


_PIN = {}
_COPY_STREAMS = {}


def h2d(x, device, dtype=None, tag=""):
    """Host array -> device tensor WITHOUT a host synchronisation: the array is copied into a
    cached pinned staging tensor (one per tag / shape / dtype) and uploaded with a non-blocking
    copy on a dedicated copy stream that the current stream then waits for.  The event that
    guards the staging tensor against being overwritten while an upload is in flight therefore
    completes as soon as that copy is done, not when the compute stream's queue has drained, so
    the host keeps running ahead of the GPU.  CUDA tensors pass through (dtype applied).
    (torch.as_tensor(numpy, device=...) uploads from pageable memory, which synchronises the
    stream: seven of those per training step were 1.3 ms of idle GPU.)"""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return x if dtype is None else x.to(dtype)
    device = torch.device(device)
    a = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    td = dtype if dtype is not None else torch.from_numpy(np.empty(0, a.dtype)).dtype
    key = (tag, str(device), tuple(a.shape), td)
    slot = _PIN.get(key)
    if slot is None:
        slot = [torch.empty(tuple(a.shape), dtype=td, pin_memory=True), None]
        _PIN[key] = slot
    if slot[1] is not None:
        slot[1].synchronize()
    slot[0].copy_(torch.from_numpy(np.ascontiguousarray(a)))
    cs = _COPY_STREAMS.get(str(device))
    if cs is None:
        cs = _COPY_STREAMS[str(device)] = torch.cuda.Stream(device)
    cur = torch.cuda.current_stream(device)
    with torch.cuda.stream(cs):
        out = slot[0].to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
    cur.wait_event(ev)
    out.record_stream(cur)
    slot[1] = ev
    return out
