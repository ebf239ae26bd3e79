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
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = f()
    for _ in range(3): g.replay()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()   # CUDA-graph replay: GPU time, not CPU launch time
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    # write-only ceiling: a plain fill of the same output bytes (torch fill kernel), same protocol
    outbuf = torch.empty(B * C * D**3 + B * D**3, dtype=torch.float32, device=dev)
    tf = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); outbuf.zero_(); e1.record(); torch.cuda.synchronize()
        tf.append(e0.elapsed_time(e1) * 1e3)
    tf.sort(); fill_us = tf[len(tf) // 2]; del outbuf
    byts = 4 * (B * P * C + 4 * B * P) + 4 * (B * C * D**3 + B * D**3)
    print(json.dumps(dict(op="avg_vox_fwd", B=B, P=P, C=C, pts=kind, us=us, min_us=ts[0], GBs=byts / us / 1e3, frac=byts / us / 1e3 / 6565.8, fill_same_bytes_us=fill_us, frac_of_fill=fill_us / us)))
