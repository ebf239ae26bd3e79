"""The sparse extractor tail alone (for ncu).  python scripts/tail_profile.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morefusion_b200 import synthetic
from morefusion_b200.contrib.singleview_3d.models import Model
dev = torch.device("cuda:0")
m = Model(n_fg_class=21, with_occupancy=True).to(dev).load_reference_weights(synthetic.init_weights(21, seed=1)).eval()
up2 = torch.randn(8, 64, 128, 128, device=dev)
pix = torch.randint(0, 65536, (8, 1000), device=dev)
with torch.no_grad():
    for _ in range(3):
        m._extractor_tail(up2, pix)
torch.cuda.synchronize()
