"""ICC accuracy/latency report on the synthetic 8-object scene (BASELINE config 4):
fused kernel vs oracle after n_iter Chainer-Adam iterations; ADD-S AUC init / oracle / ours."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from morefusion_b200 import synthetic
from morefusion_b200.contrib import IterativeCollisionCheckLink
from morefusion_b200.contrib.iterative_collision_check_link import ICCBatch
from oracle import icc as oicc, transforms as otf, metrics as om

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
out = {}
for seed, kinds in ((3, ("box",)), (4, ("box", "cylinder", "sphere"))):
    sc = synthetic.make_icc_scene(N=8, seed=seed, kinds=kinds)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
    link = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(dev)
    args = ([t(p) for p in sc["points"]], [t(s) for s in sc["sdf"]], t(sc["pitch"]), t(sc["origin"]),
            t(sc["grid_target"]), t(sc["grid_nontarget_empty"]))
    link.refine(*args, n_iter=1)   # warm-up / build
    link = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(dev)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); hist = link.refine(*args, n_iter=n_iter); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t0 = time.time()
    q_ref, t_ref, h_ref = oicc.icc_refine(sc["transform_init"], sc["points"], sc["sdf"], sc["pitch"], sc["origin"],
                                          sc["grid_target"], sc["grid_nontarget_empty"], n_iter=n_iter, sdf_offset=0.02,
                                          return_history=True)
    cpu_s = time.time() - t0
    T = otf.transformation_matrix(link.quaternion.detach().cpu().numpy(), link.translation.detach().cpu().numpy())
    T_ref = otf.transformation_matrix(q_ref, t_ref)
    rs = np.random.RandomState(0)
    surf = [synthetic.surface_points(k, h, 2000, rs).astype(np.float32) for k, h in sc["primitives"]]
    def adds(Ts):
        return np.array([om.average_distance(surf[i], sc["transform_true"][i], Ts[i])[1] for i in range(8)])
    a_init, a_ref, a_ours = adds(sc["transform_init"]), adds(T_ref), adds(T)
    dq = np.abs(T[:, :3, :3] - T_ref[:, :3, :3]).reshape(8, -1).max(1)
    out[f"seed{seed}"] = dict(
        kinds=list(kinds), per_object_max_abs_dR=[float(x) for x in dq],
        per_object_max_abs_dt=[float(x) for x in np.abs(T[:, :3, 3] - T_ref[:, :3, 3]).max(1)],
        n_iter=n_iter, gpu_ms_total=ms, gpu_us_per_iter=1e3 * ms / n_iter, oracle_cpu_s=cpu_s,
        max_abs_dt=float(np.abs(T[:, :3, 3] - T_ref[:, :3, 3]).max()),
        max_abs_dR=float(np.abs(T[:, :3, :3] - T_ref[:, :3, :3]).max()),
        loss_first=float(hist[0]), loss_last=float(hist[-1]), loss_last_oracle=float(h_ref[-1]),
        adds_auc_init=float(om.ycb_video_add_auc(a_init)), adds_auc_oracle=float(om.ycb_video_add_auc(a_ref)),
        adds_auc_ours=float(om.ycb_video_add_auc(a_ours)),
        adds_mean_mm=dict(init=float(a_init.mean() * 1e3), oracle=float(a_ref.mean() * 1e3), ours=float(a_ours.mean() * 1e3)))
# throughput: many scenes in one launch
scs = [synthetic.make_icc_scene(N=8, seed=10 + i) for i in range(4)]
for S in (1, 8, 74):
    batch = ICCBatch([scs[i % 4] for i in range(S)], sdf_offset=0.02, device=dev)
    batch.refine(n_iter=2)
    batch = ICCBatch([scs[i % 4] for i in range(S)], sdf_offset=0.02, device=dev)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); batch.refine(n_iter=30); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out[f"batch{S}"] = dict(scenes=S, iters=30, ms=ms, scene_iters_per_s=S * 30 / ms * 1e3,
                            objects_per_s_at_30_iters=S * 8 / ms * 1e3,
                            stateless_GBps=S * 30 * 2.61e6 / (ms * 1e-3) / 1e9)
print(json.dumps(out, indent=1))
open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out', 'icc_report.json'), 'w').write(json.dumps(out, indent=1))
