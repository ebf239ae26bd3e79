import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch, time
from morefusion_b200 import synthetic
from morefusion_b200.contrib.iterative_collision_check_link import ICCBatch
dev = torch.device("cuda:0")
scenes = [synthetic.make_icc_scene(N=8, seed=10 + i) for i in range(4)]
for S in (1, 8, 37, 74):
    b = ICCBatch([scenes[i % 4] for i in range(S)], sdf_offset=0.02, device=dev)
    q0, t0 = b.quaternion.clone(), b.translation.clone()
    ts = []
    for _ in range(3):
        b.quaternion.copy_(q0); b.translation.copy_(t0); b.adam_state.zero_(); b.adam_t = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.refine(n_iter=100); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(S, "scenes: ms/100it", min(ts), "scene-it/s", S * 100 / (min(ts) * 1e-3))
