import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from morefusion_b200 import synthetic
from morefusion_b200.contrib import iterative_collision_check_link as icl
dev = torch.device("cuda:0")
sc = synthetic.make_icc_scene(N=8, seed=10)
b = icl.ICCBatch([sc], sdf_offset=0.02, device=dev)
n = 12
ph = torch.zeros(1 * n * 8, dtype=torch.int64, device=dev)
aq = [icl.chainer_adam_alpha(0.01, s) for s in range(1, n + 1)]
at = [icl.chainer_adam_alpha(0.001, s) for s in range(1, n + 1)]
for _ in range(2):
    icl._run(b.prob, b.quaternion, b.translation, b.adam_state, n_iter=n, update=True, alpha_q=aq, alpha_t=at,
             voxel_threshold=2, sdf_offset=0.02, phase_ns=ph)
torch.cuda.synchronize()
t = ph.cpu().numpy().reshape(n, 8).astype(np.float64)
names = ["P0+P1 scatter", "P2 max", "P3 loss", "P4 backward", "P5 adam", "reset"]
for it in (2, 5, 9):
    row = t[it]; nxt = t[it + 1][0]
    d = [row[1] - row[0], row[2] - row[1], row[3] - row[2], row[4] - row[3], row[5] - row[4], nxt - row[5]]
    print(it, {k: round(v / 1e3, 1) for k, v in zip(names, d)}, "total", round((nxt - row[0]) / 1e3, 1))
