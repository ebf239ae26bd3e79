"""Host-side geometry helpers used at link construction time (NumPy, not on the hot path).

The reference initialises ICC parameters with trimesh.transformations.quaternion_from_matrix /
translation_from_matrix (contrib/iterative_collision_check_link.py:21-25)."""

import numpy as np


def quaternion_from_matrix(matrix):
    """Rotation part of a 4x4 (or 3x3) -> unit quaternion (w, x, y, z) with w >= 0.

    Branch-on-largest-diagonal (Shepperd) extraction: numerically stable and, up to
    round-off, the same quaternion trimesh's eigen-decomposition returns."""
    M = np.asarray(matrix, dtype=np.float64)
    m = M[:3, :3]
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0.0:
        s = np.sqrt(tr + 1.0) * 2.0
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s,
                      (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2.0
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s,
                      (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2.0
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s,
                      (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2.0
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s,
                      0.25 * s])
    q /= np.linalg.norm(q)
    if q[0] < 0.0:
        q = -q
    return q


def translation_from_matrix(matrix):
    return np.array(matrix, copy=True)[:3, 3]


# ------------------------------------------------------------------ per-frame front end (CUDA)
def pointcloud_from_depth(depth, fx, fy, cx, cy, depth_type="z"):
    """depth [H,W] float (metres, NaN = invalid) -> camera-frame points [H,W,3]
    (morefusion/geometry/pointcloud_from_depth.py:4-26) in one kernel.  numpy in -> numpy out,
    CUDA tensor in -> CUDA tensor out (no host round trip on the per-frame path)."""
    import torch
    from .. import _lib
    assert depth_type in ["z", "euclidean"], "Unexpected depth_type"
    is_np = not isinstance(depth, torch.Tensor)
    d = torch.as_tensor(np.asarray(depth), device="cuda") if is_np else depth
    assert d.dtype.is_floating_point, "depth must be float and have meter values"
    _lib.require_cuda(d)
    d = d.to(torch.float32).contiguous()
    H, W = d.shape
    pcd = torch.empty((H, W, 3), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().mf_pointcloud_from_depth(
            _lib.ptr(d), H, W, float(fx), float(fy), float(cx), float(cy),
            int(depth_type == "euclidean"), _lib.ptr(pcd), _lib.stream()), "pointcloud_from_depth")
    return pcd.cpu().numpy() if is_np else pcd


def masks_to_bboxes(masks):
    """bool masks (N,H,W) or (H,W) -> (y1, x1, y2, x2) boxes, upper bounds exclusive, zeros for
    an empty mask (morefusion/geometry/masks_to_bboxes.py:4-38).  float64 numpy for numpy input
    (as the reference), int32 CUDA tensor for CUDA input."""
    import torch
    from .. import _lib
    is_np = not isinstance(masks, torch.Tensor)
    m = torch.as_tensor(np.asarray(masks), device="cuda") if is_np else masks
    assert m.dtype == torch.bool
    assert m.dim() in [2, 3], "masks must be 2 or 3 dimensional"
    _lib.require_cuda(m)
    squeeze = m.dim() == 2
    m3 = (m[None] if squeeze else m).contiguous().view(torch.uint8)
    N, H, W = m3.shape
    out = torch.empty((N, 4), dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        _lib.check(_lib.lib().mf_masks_to_bboxes(_lib.ptr(m3), N, H, W, _lib.ptr(out), _lib.stream()),
                   "masks_to_bboxes")
    if is_np:
        out = out.cpu().numpy().astype(np.float64)
    return out[0] if squeeze else out
