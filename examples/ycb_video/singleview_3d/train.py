#!/usr/bin/env python
"""Synthetic-data twin of the reference's examples/ycb_video/singleview_3d/train.py:144-492 with
the same flags (:144-221).  One process per GPU; ``--multi-node`` expects a torchrun launch
(the reference uses ``mpirun -n 4`` + ChainerMN ``pure_nccl``, :229-233):

    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 \\
        examples/ycb_video/singleview_3d/train.py --multi-node --with-occupancy

Data: morefusion_b200.synthetic.make_rgbd_batch (YCB-shaped primitives rendered into 256 x 256
RGB-D crops; the YCB-Video datasets and pretrained weights are downloads and unavailable
offline).  Optimiser: Chainer-form Adam (alpha = --lr, :342), gradients summed over NCCL and
averaged (create_multi_node_optimizer, :344), per-rank batch = 16 // n_gpu (:361).  Snapshots
are torch state dicts (``snapshot_model_latest.pt`` / ``snapshot_trainer_latest.pt``), ``--resume``
restores both (:489-490)."""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, ROOT)

import morefusion  # noqa: E402
from morefusion_b200 import synthetic  # noqa: E402
from morefusion_b200.contrib.singleview_3d.models import training  # noqa: E402


def main():
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--multi-node", action="store_true", help="multi node")
    parser.add_argument("--gpu", type=int, default=0, help="gpu id")
    parser.add_argument("--seed", type=int, default=0, help="random seed")
    parser.add_argument("--lr", type=float, default=0.0001, help="learning rate")
    parser.add_argument("--max-epoch", type=int, default=30, help="max epoch")
    parser.add_argument("--class-ids", type=int, nargs="*", default=None, help="class id (ignored: synthetic data)")
    parser.add_argument("--pretrained-model", help="pretrained model (torch state dict)")
    parser.add_argument("--with-occupancy", action="store_true", help="with occupancy")
    parser.add_argument("--pretrained-resnet18", action="store_true", help="pretrained resnet18 (unavailable offline)")
    parser.add_argument("--resume", help="resume (output directory of a previous run)")
    parser.add_argument("--loss", choices=["add/add_s", "add", "add+occupancy", "add/add_s+occupancy"],
                        default="add/add_s", help="loss")
    parser.add_argument("--loss-scale", type=json.loads, default=None, help="loss scale")
    parser.add_argument("--out", default="logs/train_synthetic", help="output directory")
    parser.add_argument("--iters-per-epoch", type=int, default=20, help="synthetic batches per epoch")
    args = parser.parse_args()

    if args.multi_node:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank, n_gpu, device = dist.get_rank(), dist.get_world_size(), torch.device("cuda", local)
    else:
        rank, n_gpu, device = 0, 1, torch.device("cuda", args.gpu)
        torch.cuda.set_device(device)
    np.random.seed(args.seed + rank)
    torch.manual_seed(args.seed)

    model = morefusion.contrib.singleview_3d.models.Model(
        n_fg_class=21, pretrained_resnet18=args.pretrained_resnet18,
        with_occupancy=args.with_occupancy, loss=args.loss, loss_scale=args.loss_scale).to(device)
    if args.pretrained_model:
        model.load_state_dict(torch.load(args.pretrained_model, map_location=device))
    model.train()
    trainer = training.Trainer(model, alpha=args.lr)
    start_epoch = 0
    if args.resume:
        ck = torch.load(os.path.join(args.resume, "snapshot_trainer_latest.pt"), map_location=device)
        trainer.flat_p.copy_(ck["flat_p"]); trainer.flat_m.copy_(ck["flat_m"]); trainer.flat_v.copy_(ck["flat_v"])
        trainer.t, start_epoch = ck["t"], ck["epoch"]
        model._packed_ver = None
    batch_size = 16 // n_gpu
    if rank == 0:
        os.makedirs(args.out, exist_ok=True)
        json.dump(vars(args), open(os.path.join(args.out, "args.json"), "w"), indent=1)
    log = []
    for epoch in range(start_epoch, args.max_epoch):
        for it in range(args.iters_per_epoch):
            b = synthetic.make_rgbd_batch(batch_size, seed=(epoch * args.iters_per_epoch + it) * n_gpu + rank)
            kw = dict(class_id=b["class_id"], rgb=b["rgb"], pcd=b["pcd"],
                      quaternion_true=b["quaternion_true"], translation_true=b["translation_true"],
                      pitch=b["pitch"], origin=b["origin"])
            if args.with_occupancy:
                kw["grid_nontarget_empty"] = torch.as_tensor(b["grid_nontarget_empty"], device=device)
            t0 = time.time()
            loss = trainer.step(**kw)
            if rank == 0:
                model.flush_reports()
                rec = dict(epoch=epoch, iteration=trainer.t, loss=float(loss), elapsed=time.time() - t0,
                           **{k: v for k, v in model.reported.items() if k in ("add", "add_s", "add_or_add_s")})
                log.append(rec)
                print(json.dumps(rec), flush=True)
        if rank == 0:
            torch.save(model.state_dict(), os.path.join(args.out, "snapshot_model_latest.pt"))
            torch.save(dict(flat_p=trainer.flat_p, flat_m=trainer.flat_m, flat_v=trainer.flat_v,
                            t=trainer.t, epoch=epoch + 1), os.path.join(args.out, "snapshot_trainer_latest.pt"))
            json.dump(log, open(os.path.join(args.out, "log.json"), "w"), indent=1)
    if args.multi_node:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
