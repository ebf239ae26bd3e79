"""Micro-benchmark of average_voxelization_3d forward at the unit and model shapes."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morefusion_b200 as mf
mf.config.check_nan = False
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
from morefusion_b200 import synthetic
import numpy as np
for (B, P, C, kind) in [(1, 1024, 4, "uniform"), (8, 1000, 144, "uniform"), (8, 1000, 144, "surface"), (32, 1000, 144, "surface")]:
    D = 32
    if kind == "uniform":
        pts = torch.rand(B * P, 3, device=dev) * 20 + 6
    else:
        sb = synthetic.make_cnn_batch(B, P, seed=1)
        pts = torch.as_tensor(np.ascontiguousarray(sb["points"].transpose(0, 2, 1).reshape(B * P, 3)), device=dev)
    vals = torch.randn(B * P, C, device=dev)
    bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(P)
    f = lambda: mf.functions.average_voxelization_3d(vals, pts, bi, batch_size=B, origin=(0, 0, 0), pitch=1.0, dimensions=(D, D, D))
    for _ in range(5): f()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    byts = 4 * (B * P * C + 4 * B * P) + 4 * (B * C * D**3 + B * D**3)
    print(json.dumps(dict(op="avg_vox_fwd", B=B, P=P, C=C, pts=kind, us=us, min_us=ts[0], GBs=byts / us / 1e3, frac=byts / us / 1e3 / 6565.8)))
