"""Run the occupancy-map bench record alone (for ncu captures of k_map_* / quick timing).

    python scripts/mapping_profile.py [out.json]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
r = bench.bench_mapping(dev, bench.peaks(), quick=len(sys.argv) > 2)
print(json.dumps(r))
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        f.write(json.dumps(r))
