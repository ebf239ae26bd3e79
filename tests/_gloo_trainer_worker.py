"""world_size-2 gloo worker: the host-side logic of the data-parallel Trainer (flat buffers,
bucket layout, initial broadcast, bucketed gradient sum incl. buckets issued early by the
backward hook) on CPU tensors.  The CUDA parts (backward kernels, fused Adam) run in -m gpu."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morefusion_b200.contrib.singleview_3d.models import Model, training  # noqa: E402

dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + r)                      # different initial weights per rank
model = Model(n_fg_class=21, with_occupancy=True)
tr = training.Trainer(model, alpha=1e-4)
assert tr.world == 2
# 1. parameters are views into the flat buffer and were broadcast from rank 0
chk = torch.tensor([float(tr.flat_p.double().sum())], dtype=torch.float64)
both = [torch.zeros_like(chk) for _ in range(w)]
dist.all_gather(both, chk)
assert both[0].item() == both[1].item(), both
assert model.conv3.weight.data_ptr() == tr.flat_p[tr.offs["conv3.weight"][0]:].data_ptr()
assert model.conv3.weight.grad.data_ptr() == tr.flat_g[tr.offs["conv3.weight"][0]:].data_ptr()
# 2. buckets partition the flat gradient buffer in backward-completion order
spans = sorted(tr.buckets.values())
assert spans[0][0] == 0 and spans[-1][1] == tr.flat_g.numel()
assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1)), spans
assert list(tr.buckets)[:3] == ["heads", "conv4", "conv3+mlp"]
# 3. gradient sum: two buckets issued early (as the backward hook does), the rest at the end
tr.flat_g.copy_(torch.arange(tr.flat_g.numel(), dtype=torch.float32) % 97 + 1000.0 * r)
want = 2 * (torch.arange(tr.flat_g.numel(), dtype=torch.float32) % 97) + 1000.0
tr._handles = []
tr._bucket_ready("heads")
tr._bucket_ready("conv4")
tr.reduce_gradients()
assert torch.equal(tr.flat_g, want)
assert tr._handles == []
if r == 0:
    print("TRAINEROK", tr.flat_p.numel())
dist.destroy_process_group()
