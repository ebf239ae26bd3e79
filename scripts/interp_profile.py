"""interpolate_voxel_grid fwd / bwd at the two model shapes (for ncu).  python scripts/interp_profile.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import morefusion_b200 as mf
from morefusion_b200 import synthetic
dev = torch.device("cuda:0")
B, P = 8, 1000
sb = synthetic.make_cnn_batch(B, P, seed=1)
pts = torch.as_tensor(np.ascontiguousarray(sb["points"].transpose(0, 2, 1).reshape(B * P, 3)), device=dev)
bi = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(P)
for (C, D, div) in ((256, 16, 2.0), (512, 8, 4.0)):
    vox = torch.randn(B, C, D, D, D, device=dev, requires_grad=True)
    p2 = (pts / div).contiguous()
    for _ in range(2):
        y = mf.functions.interpolate_voxel_grid(vox, p2, bi)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
