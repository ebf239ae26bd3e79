"""Extra stages of __graft_entry__.smoke(): one small pass of the 3D-CNN hot path and one fused
ICC iteration on cuda:0, each checked against the oracle (the oracle is imported here only
because smoke() is one of the places allowed to use it as the checker)."""

import numpy as np
import torch


def run(dev):
    from oracle import cnn as ocnn
    from oracle import icc as oicc
    from morefusion_b200 import synthetic
    from morefusion_b200.contrib import IterativeCollisionCheckLink
    from morefusion_b200.contrib.singleview_3d.models import Model

    # ---- 3D-CNN section, 1 object
    w = ocnn.init_weights(21, seed=1)
    inp = synthetic.make_cnn_batch(1, 1000, seed=0)
    m = Model(n_fg_class=21, with_occupancy=True).to(dev).load_reference_weights(w)
    rot, trans, conf = m.forward_features(
        class_id=torch.as_tensor(inp["class_id"], device=dev),
        values=torch.as_tensor(inp["values"], device=dev),
        points=torch.as_tensor(inp["points"], device=dev), pitch=inp["pitch"], origin=inp["origin"],
        grid_nontarget_empty=torch.as_tensor(inp["grid_nontarget_empty"], device=dev))
    torch.cuda.synchronize()
    ref = ocnn.forward(w, n_fg_class=21, bf16=True, **inp)
    assert np.mean(np.abs(rot.cpu().numpy() - ref["rot"])) < 5e-3, "CNN rot mismatch"
    assert np.mean(np.abs(conf.cpu().numpy() - ref["conf"])) < 5e-3, "CNN conf mismatch"
    assert any(k[0] == "tc" for k in m.launch_log), "tcgen05 GEMM path did not run"

    # ---- fused ICC: loss + gradient of one iteration, 2 objects, 16^3 grids
    rs = np.random.RandomState(0)
    D = 16
    pts, sdf, T0, origin, gt, gne = [], [], [], [], [], []
    pitch = np.array([0.0063, 0.0087], np.float32)
    centers = np.array([[0.0, 0.0, 0.6], [0.05, 0.01, 0.61]], np.float32)
    for i in range(2):
        r = pitch[i] * D * 0.3
        ax = np.arange(-r, r + 1e-9, pitch[i])
        g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
        d = np.linalg.norm(g, axis=1)
        pts.append(g[d <= r].astype(np.float32))
        sdf.append((r - d[d <= r]).astype(np.float32))
        T = np.eye(4, dtype=np.float32)
        T[:3, 3] = centers[i] + rs.normal(0, 0.003, 3)
        T0.append(T)
        origin.append(centers[i] - pitch[i] * (D / 2.0 - 0.5))
        gt.append((rs.uniform(size=(D, D, D)) < 0.05).astype(np.float32))
        gne.append((rs.uniform(size=(D, D, D)) < 0.5).astype(np.float32))
    origin, gt, gne = np.stack(origin).astype(np.float32), np.stack(gt), np.stack(gne)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    link = IterativeCollisionCheckLink(np.stack(T0), voxel_dim=D, sdf_offset=0.02).to(dev)
    loss = link([t(p) for p in pts], [t(s) for s in sdf], t(pitch), t(origin), t(gt), t(gne))
    loss.backward()
    r = oicc.icc_forward_backward(link.quaternion.detach().cpu().numpy(),
                                  link.translation.detach().cpu().numpy(), pts, sdf, pitch, origin,
                                  gt, gne, voxel_dim=D, sdf_offset=0.02)
    assert abs(float(loss.detach()) - float(r["loss"])) < 1e-4 * max(1.0, abs(float(r["loss"]))), "ICC loss"
    gt_ = link.translation.grad.cpu().numpy()
    assert np.abs(gt_ - r["gt"]).max() < 1e-3 * max(1.0, np.abs(r["gt"]).max()), "ICC grad"
