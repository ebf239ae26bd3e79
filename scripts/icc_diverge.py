"""Diagnostic: run the ICC driver loop on the GPU and in the oracle from the same state and report
the first iteration at which gradients / parameters stop being bit-identical."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import icc as oicc
from oracle.ref_harness import gen_icc_closed_loop as gen
from morefusion_b200.contrib import IterativeCollisionCheckLink

name = sys.argv[1] if len(sys.argv) > 1 else "ref3"
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dev = torch.device("cuda:0")
g = np.load(os.path.join(gen.OUT, f"icc_closed_loop_{name}.npz"))
if name == "ref3":
    off = np.r_[0, np.cumsum(g["sizes"])]
    sc = dict(points=[g["points"][off[i]:off[i + 1]] for i in range(3)],
              sdf=[g["sdf"][off[i]:off[i + 1]] for i in range(3)], pitch=g["pitch"], origin=g["origin"],
              grid_target=g["grid_target"].astype(np.float32),
              grid_nontarget_empty=g["grid_nontarget_empty"].astype(np.float32),
              transform_init=g["transform_init"])
else:
    sc = gen.scene(name)
t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
args = ([t(p) for p in sc["points"]], [t(x) for x in sc["sdf"]], t(sc["pitch"]), t(sc["origin"]),
        t(sc["grid_target"]), t(sc["grid_nontarget_empty"]))
np_args = (sc["points"], sc["sdf"], sc["pitch"], sc["origin"], sc["grid_target"], sc["grid_nontarget_empty"])
link = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(dev)
q = link.quaternion.detach().cpu().numpy().copy()
tr = link.translation.detach().cpu().numpy().copy()
oq, ot = oicc.ChainerAdam(q.shape, 0.01), oicc.ChainerAdam(tr.shape, 0.001)
first = None
for it in range(n_iter):
    # gradients at the ORACLE's state, both sides
    with torch.no_grad():
        link.quaternion.copy_(t(q)); link.translation.copy_(t(tr))
    link.zero_grad()
    loss = link(*args); loss.backward()
    r = oicc.icc_forward_backward(q, tr, *np_args, sdf_offset=0.02)
    gq, gt = link.quaternion.grad.cpu().numpy(), link.translation.grad.cpu().numpy()
    same = np.array_equal(gq, r["gq"]) and np.array_equal(gt, r["gt"]) and float(loss) == float(r["loss"])
    if not same:
        d = dict(it=it, loss_gpu=float(loss), loss_oracle=float(r["loss"]),
                 dgq=np.abs(gq - r["gq"]).max().item(), dgt=np.abs(gt - r["gt"]).max().item(),
                 gq_scale=np.abs(r["gq"]).max().item(), gt_scale=np.abs(r["gt"]).max().item(),
                 where_q=np.argwhere(gq != r["gq"]).tolist(), where_t=np.argwhere(gt != r["gt"]).tolist())
        print("MISMATCH", json.dumps(d))
        if first is None:
            first = it
            np.savez("gpurun_out/icc_diverge_state.npz", q=q, t=tr, gq_gpu=gq, gt_gpu=gt, gq=r["gq"], gt=r["gt"], gsum=r["gsum"])
    oq.update(q, r["gq"]); ot.update(tr, r["gt"])
print("first mismatch:", first)
# closed loop on the GPU for the same number of iterations vs the oracle's final state
link2 = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(dev)
link2.refine(*args, n_iter=n_iter)
print("closed loop max|dq|", np.abs(link2.quaternion.detach().cpu().numpy() - q).max(),
      "max|dt|", np.abs(link2.translation.detach().cpu().numpy() - tr).max())
