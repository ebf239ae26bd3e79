"""Where does the training step spend its time?  torch.profiler over a few Trainer steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import morefusion_b200 as mf
from morefusion_b200 import synthetic
from morefusion_b200.contrib.singleview_3d.models import Model, training
mf.config.check_nan = False
dev = torch.device("cuda:0")
Bl = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = Model(n_fg_class=21, with_occupancy=True).to(dev).load_reference_weights(synthetic.init_weights(21, seed=1)).train()
tr = training.Trainer(model, alpha=1e-4)
b = synthetic.make_cnn_batch(Bl, 1000, seed=0)
t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
d = {k: t(v) for k, v in b.items()}
rs = np.random.RandomState(0)
q = rs.normal(size=(Bl, 4)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
tt = (b["points"] * b["pitch"][:, None, None] + b["origin"][:, :, None]).mean(axis=2).astype(np.float32)
model.predict = lambda **kw: training.forward_features_with_grad(
    model, class_id=b["class_id"], values=d["values"], points=d["points"], pitch=d["pitch"],
    origin=d["origin"], grid_nontarget_empty=d["grid_nontarget_empty"])
step = lambda: tr.step(class_id=b["class_id"], rgb=None, pcd=None, quaternion_true=q, translation_true=tt)
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.time()
for _ in range(5): step()
torch.cuda.synchronize()
print("wall ms/step", (time.time() - t0) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60))
