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

    # ---- one training step of the 3-D section (forward + CUDA backward + fused Adam), 1 object:
    # the conv3 weight gradient against the differentiable oracle
    from oracle import cnn_train as ct
    from morefusion_b200.contrib.singleview_3d.models import training
    import morefusion_b200 as mf
    m.train()
    q_true = np.array([[1, 0, 0, 0]], np.float32)
    cam = inp["points"] * inp["pitch"][:, None, None] + inp["origin"][:, :, None]
    t_true = cam.mean(axis=2).astype(np.float32)
    cad = [synthetic.SyntheticYCBModels().get_pcd(int(inp["class_id"][0]))[:500]]
    want_loss, want, _ = ct.loss_and_grads(w, inp, quaternion_true=q_true, translation_true=t_true,
                                           cad_points=cad, symmetric=[False])
    rot, trans, conf = training.forward_features_with_grad(
        m, class_id=inp["class_id"], values=t(inp["values"]), points=t(inp["points"]),
        pitch=t(inp["pitch"]), origin=t(inp["origin"]),
        grid_nontarget_empty=t(inp["grid_nontarget_empty"]))
    F = mf.functions
    T_pred = F.transformation_matrix(rot[0], trans[0])
    T_true = F.transformation_matrix(t(q_true[0]), t(t_true[0]))
    add = F.average_distance(t(cad[0]), T_true, T_pred, symmetric=False)
    loss = torch.mean(add * conf[0] - 0.015 * torch.log(conf[0]))
    loss.backward()
    assert abs(float(loss.detach()) - want_loss) < 3e-2 * abs(want_loss) + 1e-4, "training loss"
    g, r = m.conv3.weight.grad.float().cpu().numpy().ravel(), want["conv3/W"].ravel()
    cos = float(g @ r / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
    assert cos > 0.995, f"conv3 weight gradient (tcgen05 MN-major wgrad) cos={cos}"
    # ---- occupancy-grid producer: one small scan + query vs the OctoMap restatement
    from oracle import octomap as oc
    from test_mapping_emu import make_scene
    pcd, fg = make_scene(4)
    ours, ref = mf.contrib.MultiInstanceOctreeMapping(capacity=1 << 14), oc.MultiInstanceOctreeMapping()
    for mp in (ours, ref):
        mp.initialize(1, pitch=0.008)
        mp.initialize(0, pitch=0.02)
        mp.integrate(1, fg, pcd)
        mp.integrate(0, ~fg, pcd)
    org = np.nanmedian(pcd[fg], axis=0) - 7.5 * 0.008
    got = ours.get_target_grids(1, dimensions=(16, 16, 16), pitch=0.008, origin=org)
    want = ref.get_target_grids(1, dimensions=(16, 16, 16), pitch=0.008, origin=org)
    for a, b in zip(got, want):
        assert np.array_equal(a > 0, b > 0) and np.allclose(a, b, rtol=0, atol=1e-7), "occupancy map"
