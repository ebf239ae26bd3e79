"""In-graph timeline of the inference step: replay the Runner's CUDA graph under torch.profiler
(CUPTI kernel activities, no replay/serialisation) and print each kernel's start offset and
duration relative to the step start, averaged over the profiled steps (L2 flushed between)."""
import json, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import morefusion_b200 as mf
from morefusion_b200 import synthetic
from morefusion_b200.contrib.singleview_3d.models import Model
from torch.profiler import profile, ProfilerActivity

mf.config.check_nan = False
dev = torch.device("cuda:0")
B, P = 8, 1000
m = Model(n_fg_class=21, with_occupancy=True).to(dev).load_reference_weights(synthetic.init_weights(21, seed=1))
FLUSH = "write"
for a in sys.argv[1:]:
    if a.startswith("flush="):
        FLUSH = a.split("=")[1]
    elif a.startswith("--"):
        pass
    elif "=" in a:
        k, v = a.split("="); setattr(m, k, json.loads(v))
r = m.make_runner(B, P, dev, graph=True)
r.load_host(synthetic.make_cnn_batch(B, P, seed=0))
fbuf = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
rbuf = torch.zeros(48 << 20, dtype=torch.int32, device=dev)


class flush:
    @staticmethod
    def zero_():
        if FLUSH in ("write", "writeread"):
            fbuf.zero_()
        if FLUSH in ("read", "writeread"):
            rbuf.max()          # 192 MiB read: L2 left full of clean lines


E2E = "--e2e" in sys.argv
run = r.run_e2e if E2E else r.run
for _ in range(5):
    flush.zero_(); run()
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        flush.zero_(); run()
        if E2E:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into steps at the flush kernels (or evenly when there is no flush)
is_sep = lambda e: "FillFunctor" in e.name or "reduce_kernel" in e.name   # noqa: E731
steps, cur = [], None
if FLUSH == "none":
    ks = list(evs)
    n = len(ks) // N
    steps = [ks[i * n:(i + 1) * n] for i in range(N)]
else:
    for e in evs:
        if is_sep(e):
            if cur:
                steps.append(cur)
            cur = []
        elif cur is not None:
            cur.append(e)
    if cur:
        steps.append(cur)
steps = [s for s in steps if len(s) == len(steps[-1])]
if "--raw" in sys.argv:
    for s in steps[3:5]:
        t0 = s[0].time_range.start
        print("--- one step")
        for e in s:
            print(f"{e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:7.1f} ={e.time_range.end - t0:7.1f}  {e.name[:70]}")
    sys.exit(0)
agg = collections.OrderedDict()
tot = []
for s in steps:
    t0 = s[0].time_range.start
    tot.append(max(e.time_range.end for e in s) - t0)
    for i, e in enumerate(s):
        a = agg.setdefault(i, [e.name[:70], 0.0, 0.0])
        a[1] += (e.time_range.start - t0) / len(steps); a[2] += (e.time_range.end - e.time_range.start) / len(steps)
print(f"flush={FLUSH} steps={len(steps)}  span_us={sum(tot)/len(tot):.1f}   (start, +dur, end)")
for i, (n, st, du) in agg.items():
    print(f"{st:8.1f} +{du:7.1f} ={st+du:7.1f}  {n}")
