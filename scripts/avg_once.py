"""Run average_voxelization_3d once per shape, eagerly (for `ncu` launch lists)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import morefusion_b200 as mf

dev = torch.device("cuda:0")
mf.config.check_nan = False
for B, kind in ((8, "uniform"), (8, "surface"), (32, "surface")):
    rs = np.random.RandomState(0)
    P, C, D = 1000, 144, 32
    if kind == "uniform":
        pts = rs.uniform(0, D - 1, (B * P, 3)).astype(np.float32)
    else:
        d = rs.normal(size=(B * P, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        pts = (16 + d * rs.uniform(6, 11, (B * P, 1))).astype(np.float32)
    vals = torch.tensor(rs.normal(size=(B * P, C)).astype(np.float32), device=dev)
    bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(P)
    pts = torch.tensor(pts, device=dev)
    for _ in range(2):
        y = mf.functions.average_voxelization_3d(vals, pts, bi, batch_size=B, origin=(0, 0, 0),
                                                 pitch=1.0, dimensions=(D, D, D))
    torch.cuda.synchronize()
    del y
