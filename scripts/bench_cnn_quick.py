import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
import morefusion_b200 as mf
from morefusion_b200.contrib.singleview_3d.models import Model
from oracle import cnn as ocnn
from test_cnn_gpu import make_inputs
mf.config.check_nan = False
dev = torch.device("cuda:0")
B = 8
w = ocnn.init_weights(21, seed=1)
inp = make_inputs(B)
m = Model(n_fg_class=21, with_occupancy=True).to(dev).load_reference_weights(w)
m.use_tensor_cores = "--simt" not in sys.argv
args = dict(class_id=torch.as_tensor(inp["class_id"], device=dev), values=torch.as_tensor(inp["values"], device=dev),
            points=torch.as_tensor(inp["points"], device=dev), pitch=torch.as_tensor(inp["pitch"], device=dev),
            origin=torch.as_tensor(inp["origin"], device=dev),
            grid_nontarget_empty=torch.as_tensor(inp["grid_nontarget_empty"], device=dev))
for _ in range(3): m.forward_features(**args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5): m.forward_features(**args)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps(dict(B=B, ms_per_step=ms, objects_per_s=B / ms * 1e3, tc=m.use_tensor_cores)))
