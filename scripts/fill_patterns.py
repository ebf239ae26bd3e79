"""Store-pattern probe: the same 152 MB of zeros written five ways (mf_debug_fill_probe)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morefusion_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda:0")
B, planes, V = 8, 145, 32768
out = torch.empty(B * planes * V, dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in (0, 1, 2, 3, 4, 0, 1):
    ts = []
    for _ in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        rc = L.mf_debug_fill_probe(_lib.ptr(out), B, planes, V, mode, _lib.stream())
        e1.record(); torch.cuda.synchronize()
        assert rc == 0
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(json.dumps(dict(mode=mode, us=ts[len(ts) // 2], min_us=ts[0], GBs=out.numel() * 4 / ts[len(ts) // 2] / 1e3)))
