import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from morefusion_b200 import synthetic
from morefusion_b200.contrib.iterative_collision_check_link import ICCBatch
dev = torch.device("cuda:0")
scs = [synthetic.make_icc_scene(N=8, seed=10 + i) for i in range(4)]
for S in (1, 4, 16):
    for G in (0, 16, 37, 74, 148, 296):
        if S * max(G, 1) > 296 and G != 0: continue
        batch = ICCBatch([scs[i % 4] for i in range(S)], sdf_offset=0.02, device=dev)
        batch.group_size = G
        batch.refine(n_iter=2)
        batch = ICCBatch([scs[i % 4] for i in range(S)], sdf_offset=0.02, device=dev)
        batch.group_size = G
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); batch.refine(n_iter=30); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps(dict(scenes=S, G=G, us_per_iter=round(ms / 30 * 1e3, 1), scene_iters_per_s=round(S * 30 / ms * 1e3))), flush=True)
