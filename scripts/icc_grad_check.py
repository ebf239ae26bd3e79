"""Per-iteration gradient agreement along the ORACLE trajectory (isolates single-step parity
from Adam's amplification of round-off)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from morefusion_b200 import synthetic
from morefusion_b200.contrib import IterativeCollisionCheckLink
from oracle import icc as oicc, transforms as otf
dev = torch.device("cuda:0")
sc = synthetic.make_icc_scene(N=8, seed=3, kinds=("box",))
t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)
args = ([t(p) for p in sc["points"]], [t(s) for s in sc["sdf"]], t(sc["pitch"]), t(sc["origin"]),
        t(sc["grid_target"]), t(sc["grid_nontarget_empty"]))
link = IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(dev)
q = np.stack([otf.quaternion_from_matrix(T) for T in sc["transform_init"]]).astype(np.float32)
tr = sc["transform_init"][:, :3, 3].astype(np.float32).copy()
oq, ot = oicc.ChainerAdam(q.shape, 0.01), oicc.ChainerAdam(tr.shape, 0.001)
for it in range(8):
    r = oicc.icc_forward_backward(q, tr, sc["points"], sc["sdf"], sc["pitch"], sc["origin"], sc["grid_target"],
                                  sc["grid_nontarget_empty"], sdf_offset=0.02)
    with torch.no_grad():
        link.quaternion.copy_(t(q)); link.translation.copy_(t(tr))
    link.zero_grad()
    loss = link(*args); loss.backward()
    gq, gt = link.quaternion.grad.cpu().numpy(), link.translation.grad.cpu().numpy()
    eq = np.abs(gq - r["gq"]) / (np.abs(r["gq"]).max(1, keepdims=True) + 1e-12)
    et = np.abs(gt - r["gt"]) / (np.abs(r["gt"]).max(1, keepdims=True) + 1e-12)
    print(it, "loss", float(loss), float(r["loss"]), "max rel gq err per obj", ["%.1e" % x for x in eq.max(1)],
          "gt", ["%.1e" % x for x in et.max(1)])
    oq.update(q, r["gq"]); ot.update(tr, r["gt"])
