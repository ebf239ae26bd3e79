import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morefusion_b200 import synthetic
from morefusion_b200.contrib.iterative_collision_check_link import ICCBatch
dev = torch.device("cuda:0")
sc = synthetic.make_icc_scene(N=8, seed=10)
b = ICCBatch([sc], sdf_offset=0.02, device=dev)
b.refine(n_iter=10)
torch.cuda.synchronize()
