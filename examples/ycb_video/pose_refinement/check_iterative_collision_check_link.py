#!/usr/bin/env python
"""Synthetic-data twin of the reference's
examples/ycb_video/pose_refinement/check_iterative_collision_check_link.py:13-79.

Same driver: IterativeCollisionCheckLink(transform, sdf_offset=0.02), Adam(alpha=0.01) with the
translation alpha scaled by 0.1, 100 x (forward, backward, update).  The reference needs the YCB
model download (models.get_sdf) and an OpenGL viewer; here the scene is either the reference's own
committed 3-object scene with analytic box SDFs (tests/golden/icc_closed_loop_ref3.npz) or a
synthetic N-object scene, and the result is printed instead of drawn.

    python examples/ycb_video/pose_refinement/check_iterative_collision_check_link.py [--scene ref3|synthetic]
        [--n-objects 8] [--seed 3] [--iterations 100] [--fused]
"""

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)

import morefusion  # noqa: E402  (alias of morefusion_b200)
from morefusion_b200 import synthetic  # noqa: E402
from morefusion_b200.optimizers import ChainerAdam  # noqa: E402


def get_scene(args):
    if args.scene == "ref3":
        g = np.load(os.path.join(ROOT, "tests", "golden", "icc_closed_loop_ref3.npz"))
        off = np.r_[0, np.cumsum(g["sizes"])]
        return dict(points=[g["points"][off[i]:off[i + 1]] for i in range(3)],
                    sdf=[g["sdf"][off[i]:off[i + 1]] for i in range(3)], pitch=g["pitch"],
                    origin=g["origin"], grid_target=g["grid_target"].astype(np.float32),
                    grid_nontarget_empty=g["grid_nontarget_empty"].astype(np.float32),
                    transform_init=g["transform_init"])
    return synthetic.make_icc_scene(N=args.n_objects, seed=args.seed)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    ap.add_argument("--scene", choices=["ref3", "synthetic"], default="ref3")
    ap.add_argument("--n-objects", type=int, default=8)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--iterations", type=int, default=100)
    ap.add_argument("--fused", action="store_true",
                    help="all iterations in one persistent kernel (link.refine) instead of the "
                         "reference's python loop")
    args = ap.parse_args()
    dev = torch.device("cuda")
    sc = get_scene(args)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    points, sdf = [t(p) for p in sc["points"]], [t(s) for s in sc["sdf"]]
    pitch, origin = t(sc["pitch"]), t(sc["origin"])
    grid_target, grid_nontarget_empty = t(sc["grid_target"]), t(sc["grid_nontarget_empty"])

    link = morefusion.contrib.IterativeCollisionCheckLink(sc["transform_init"], sdf_offset=0.02).to(dev)
    t0 = time.time()
    if args.fused:
        hist = link.refine(points, sdf, pitch, origin, grid_target, grid_nontarget_empty,
                           n_iter=args.iterations).cpu().numpy()
    else:
        optimizer = ChainerAdam([dict(params=[link.quaternion], alpha=0.01),
                                 dict(params=[link.translation], alpha=0.01 * 0.1)])
        hist = []
        for i in range(args.iterations):
            loss = link(points, sdf, pitch, origin, grid_target, grid_nontarget_empty)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            optimizer.step()
            hist.append(float(loss))
    torch.cuda.synchronize()
    dt = time.time() - t0
    T = morefusion.functions.transformation_matrix(link.quaternion, link.translation)
    print(f"{args.iterations} iterations in {dt * 1e3:.1f} ms; loss {hist[0]:.5f} -> {hist[-1]:.5f}")
    if "transform_true" in sc:
        err = np.linalg.norm(T[:, :3, 3].detach().cpu().numpy() - sc["transform_true"][:, :3, 3], axis=1)
        err0 = np.linalg.norm(sc["transform_init"][:, :3, 3] - sc["transform_true"][:, :3, 3], axis=1)
        print("translation error (mm): init", np.round(err0 * 1e3, 2), "-> refined", np.round(err * 1e3, 2))


if __name__ == "__main__":
    main()
