"""One short ICC launch for `ncu --set full -k regex:k_icc_run` (S scenes x n iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morefusion_b200 import synthetic
from morefusion_b200.contrib.iterative_collision_check_link import ICCBatch
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
scenes = [synthetic.make_icc_scene(N=8, seed=10 + i) for i in range(min(S, 4))]
b = ICCBatch([scenes[i % len(scenes)] for i in range(S)], sdf_offset=0.02, device=dev)
b.refine(n_iter=n)
torch.cuda.synchronize()
