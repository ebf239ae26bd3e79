import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dist.init_process_group("gloo")
r = dist.get_rank()
v = bench.max_over_ranks(10.0 + r, 2, torch.device("cpu"))
bench.barrier(2)
assert v == 11.0, v
if r == 0:
    print("MAXOK")
dist.destroy_process_group()
