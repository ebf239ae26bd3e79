"""Single-scene ICC latency against the number of CTAs that share the scene (group barrier cost
against parallel width).  python scripts/icc_group_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morefusion_b200 import synthetic  # noqa: E402
from morefusion_b200.contrib.iterative_collision_check_link import ICCBatch  # noqa: E402

dev = torch.device("cuda:0")
sc = synthetic.make_icc_scene(N=8, seed=10)
for S in (1, 2):
    for G in (0, 592, 444, 296, 222, 148, 74):
        batch = ICCBatch([sc] * S, sdf_offset=0.02, device=dev)
        batch.group_size = G
        q0, t0 = batch.quaternion.clone(), batch.translation.clone()
        ts = []
        for _ in range(4):
            batch.quaternion.copy_(q0); batch.translation.copy_(t0)
            batch.adam_state.zero_(); batch.adam_t = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); batch.refine(n_iter=100); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"scenes {S} group_size {G:4d}: {min(ts[1:]) * 10:.1f} us/iteration", flush=True)
