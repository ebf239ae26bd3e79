"""Training step of the 3-D section of singleview_3d.Model on the sm_100a kernels.

What chainer's reverse pass + ChainerMN do for
  morefusion/contrib/singleview_3d/models/model.py:93-141 (_extract), :239-273 (heads / pose),
  examples/ycb_video/singleview_3d/train.py:229-233 (pure_nccl communicator), :342-344 (Adam +
  multi-node optimizer), :361 (per-rank batch = 16 / n_gpu)
is here:
  * ``forward_features_with_grad``: a torch.autograd.Function whose forward is the model's fixed
    CUDA launch sequence (cnn.cu / conv3d_tc.cu) and whose backward is the launch sequence below
    (train.cu / gemm_train.cu): head GEMM gradients (tcgen05: dX on the forward kernel with
    transposed weights, dW with MN-major operands), trilinear scatter, ReLU mask + padding, conv4
    / conv3 weight gradients (MN-major implicit GEMM over the s2d input) and input gradients
    (8 parity GEMMs over the padded dY), voxelisation backward, per-point MLP backward.  The two
    thin occupancy stencils (0.2 % of the FLOPs) take their gradient from torch's conv3d.
  * ``Trainer``: parameters / gradients / Adam moments live in flat fp32 buffers (the parameters
    of the model are views into them); gradients are all-reduced over NCCL in buckets that are
    issued as soon as the backward has finished writing them, overlapping the remaining backward
    kernels; one fused kernel applies the 1/world scale and the Chainer-form Adam update.
"""

import numpy as np
import torch

from .... import _lib
from ....functions.geometry import _util
from . import model as M

HEADS = ("rot", "trans", "conf")


# ------------------------------------------------------------------ parameter order
def flat_param_order(model):
    """Names of the 3-D section's parameters in the order the backward pass finishes them (the
    all-reduce buckets follow it) with same-shape groups contiguous (grouped GEMM outputs)."""
    names = []
    for layer in (4, 3, 2, 1):
        names += [f"conv{layer}_{h}.weight" for h in HEADS]
        names += [f"conv{layer}_{h}.bias" for h in HEADS]
    names += ["conv4.weight", "conv4.bias", "conv3.weight", "conv3.bias"]
    if model._with_occupancy:
        names += ["conv2_occ.weight", "conv2_occ.bias", "conv1_occ.weight", "conv1_occ.bias"]
    names += ["conv2_rgb.weight", "conv2_rgb.bias", "conv2_pcd.weight", "conv2_pcd.bias",
              "conv1_rgb.weight", "conv1_rgb.bias", "conv1_pcd.weight", "conv1_pcd.bias"]
    return names


def pack_dgrad_weight(W):
    """OIDHW [Co,Ci,4,4,4] -> bf16 [8][Ci][8*Co]: Wd[r][ci][a*Co+co] = W[co][ci][2a+r] per axis
    (input voxel parity r, tap a): the B operand of the 8 parity GEMMs of the input gradient."""
    Co, Ci = W.shape[:2]
    W = W.reshape(Co, Ci, 2, 2, 2, 2, 2, 2)                # co ci ad rd ah rh aw rw
    W = W.permute(3, 5, 7, 1, 2, 4, 6, 0)                  # rd rh rw ci ad ah aw co
    return W.reshape(8, Ci, 8 * Co).to(torch.bfloat16).contiguous()


def unpack_conv_k4s2_grad(G, Co, Ci):
    """[Co, 64*Ci] in the forward's packed K order (pack_conv_k4s2_weight) -> OIDHW."""
    G = G.reshape(Co, 2, 2, 2, 2, 2, 2, Ci)                # co ad ah aw rd rh rw ci
    G = G.permute(0, 7, 1, 4, 2, 5, 3, 6)                  # co ci ad rd ah rh aw rw
    return G.reshape(Co, Ci, 4, 4, 4)


def _train_pack(model):
    """bf16 operands the backward GEMMs need besides the forward's packed weights."""
    ver = tuple(p._version for p in model.parameters())
    tp = getattr(model, "_train_packed", None)
    if tp is not None and tp["ver"] == ver:
        return tp
    bf = torch.bfloat16
    p = dict(ver=ver)
    W1 = torch.cat([getattr(model, f"conv1_{h}").weight.detach().reshape(640, 984) for h in HEADS], 0)
    W1t = torch.zeros(992, 1920, dtype=bf, device=W1.device)       # N padded 984 -> 992 (x32)
    W1t[:984] = W1.t().to(bf)
    p["head1/Wt"] = W1t
    for h in HEADS:
        for layer in (2, 3):
            m = getattr(model, f"conv{layer}_{h}")
            p[f"conv{layer}_{h}/Wt"] = m.weight.detach().reshape(m.weight.shape[0], -1).t().to(bf).contiguous()
    p["conv3/Wd"] = pack_dgrad_weight(model.conv3.weight.detach().float())
    p["conv4/Wd"] = pack_dgrad_weight(model.conv4.weight.detach().float())
    for n in ("conv1_rgb", "conv1_pcd", "conv2_rgb", "conv2_pcd"):
        m = getattr(model, n)
        p[n + "/W"] = m.weight.detach().reshape(m.weight.shape[0], -1).float().contiguous()   # [out][in]
    model._train_packed = p
    return p


def _train_buffers(model, B, P, dev):
    key = ("train", B, P, dev)
    if key in model._wbufs:
        return model._wbufs[key]
    bf, f32 = torch.bfloat16, torch.float32
    z = lambda *s, dt=bf: torch.zeros(*s, dtype=dt, device=dev)     # noqa: E731
    NP = B * P
    Ct = 144 + (16 if model._with_occupancy else 0)
    b = dict(
        raw8=z(NP, 8, dt=f32), dhd3=z(NP, 384), dhd2=z(NP, 768), dhd1=z(NP, 1920),
        dfeat=z(NP, M.FEAT_LD), dgrid4=z(B, 8 ** 3, 512, dt=f32), dgrid3=z(B, 16 ** 3, 256, dt=f32),
        dY4p=z(B, 10, 10, 10, 512), dY3p=z(B, 18, 18, 18, 256),
        dx3=z(8, B * 4096, Ct), dfeat2=z(NP, 144, dt=f32),
        gw4=z(512, 64 * 256, dt=f32), gw3=z(256, 64 * Ct, dt=f32))
    model._wbufs[key] = b
    return b


# ------------------------------------------------------------------ autograd function
class _PoseNet3D(torch.autograd.Function):
    """values -> (rot, trans, conf) through the CUDA forward; backward = CUDA launch sequence.
    The activations live in the model's persistent work buffers: run backward before the next
    forward of the same batch shape (the usual training loop)."""

    @staticmethod
    def forward(ctx, model, st, values, *params):
        B, _, P = values.shape
        dev = values.device
        st = dict(st, values=values.detach().contiguous().float())
        tb = _train_buffers(model, B, P, dev)
        out = dict(rot=torch.empty((B, P, 4), dtype=torch.float32, device=dev),
                   trans=torch.empty((B, P, 3), dtype=torch.float32, device=dev),
                   conf=torch.empty((B, P), dtype=torch.float32, device=dev))
        model._raw8 = tb["raw8"]
        try:
            model._stage_pre(st)
            model._stage_conv3(st)
            model._stage_post(st, out)
        finally:
            model._raw8 = None
        ctx.model, ctx.st, ctx.shape = model, st, (B, P)
        ctx.names = flat_param_order(model)
        ctx.needs_values = values.requires_grad
        return out["rot"], out["trans"], out["conf"]

    @staticmethod
    def backward(ctx, g_rot, g_trans, g_conf):
        model, st = ctx.model, ctx.st
        grads = backward_3d(model, st, g_rot.contiguous().float(), g_trans.contiguous().float(),
                            g_conf.contiguous().float(), ctx.needs_values)
        dvalues = grads.pop("values", None)
        if getattr(model, "_train_grad_views", None) is not None:
            # Trainer mode: the gradients were written straight into the flat buffer that the
            # parameters' .grad views alias -- returning them would make autograd add them again
            return (None, None, dvalues) + (None,) * len(ctx.names)
        return (None, None, dvalues) + tuple(grads[n] for n in ctx.names)


def forward_features_with_grad(model, *, class_id, values, points, pitch, origin,
                               grid_nontarget_empty=None):
    _lib.require_cuda(values, points)
    dev = values.device
    if not model.fused_head4 or not model.fused_voxelize or not model.use_tensor_cores:
        raise RuntimeError("the training step runs on the fused tensor-core forward")
    st = dict(points=points.detach().contiguous().float(),
              class_id=_util.h2d(class_id, dev, torch.int32, "train/class_id").contiguous(),
              pitch=torch.as_tensor(pitch, dtype=torch.float32, device=dev).contiguous(),
              origin=torch.as_tensor(origin, dtype=torch.float32, device=dev).contiguous(),
              gne=None if grid_nontarget_empty is None
              else torch.as_tensor(grid_nontarget_empty, device=dev))
    named = dict(model.named_parameters())
    params = [named[n] for n in flat_param_order(model)]
    return _PoseNet3D.apply(model, st, values, *params)


def _out_grads(model):
    """fp32 gradient destinations: the Trainer's views into the flat gradient buffer (already
    zeroed) if there is one, else fresh zero tensors."""
    tg = getattr(model, "_train_grad_views", None)
    if tg is not None:
        return dict(tg)
    named = dict(model.named_parameters())
    return {n: torch.zeros_like(named[n], dtype=torch.float32) for n in flat_param_order(model)}


def backward_3d(model, st, g_rot, g_trans, g_conf, needs_values=False, on_bucket=None):
    """Launch sequence of the backward pass.  Returns {param name: gradient} (+ 'values')."""
    L = _lib.lib()
    s = _lib.stream
    ptr = _lib.ptr
    B, _, P = st["values"].shape
    dev = st["values"].device
    NP = B * P
    nfg = model._n_fg_class
    D = model._voxel_dim
    Ct = 144 + (16 if model._with_occupancy else 0)
    w = model._packed
    tw = _train_pack(model)
    buf = model._work_buffers(B, P, dev)
    tb = _train_buffers(model, B, P, dev)
    G = _out_grads(model)
    on_bucket = on_bucket or getattr(model, "_on_bucket", None)
    chk = _lib.check
    with torch.cuda.device(dev):
        tb["dgrid4"].zero_()
        tb["dgrid3"].zero_()
        # ---- layer 4 of the heads + pose epilogue
        chk(L.mf_train_head4_bwd(
            ptr(g_rot), ptr(g_trans), ptr(g_conf), ptr(tb["raw8"]), ptr(buf["hd3"]), 384,
            ptr(w["conv4_rot/W"]), ptr(w["conv4_trans/W"]), ptr(w["conv4_conf/W"]),
            ptr(st["class_id"]), ptr(st["pitch"]), B, P, nfg, ptr(tb["dhd3"]),
            ptr(G["conv4_rot.weight"]), ptr(G["conv4_rot.bias"]), ptr(G["conv4_trans.weight"]),
            ptr(G["conv4_trans.bias"]), ptr(G["conv4_conf.weight"]), ptr(G["conv4_conf.bias"]), s()),
            "head4_bwd")
        # ---- layers 3, 2: dW (MN-major tcgen05), db, dX (forward GEMM on W^T), ReLU mask
        for layer, dz, x, n_out, k_in, dx, act in (
                (3, tb["dhd3"], buf["hd2"], 128, 256, tb["dhd2"], buf["hd2"]),
                (2, tb["dhd2"], buf["hd1"], 256, 640, tb["dhd1"], buf["hd1"])):
            gw = [G[f"conv{layer}_{h}.weight"] for h in HEADS]
            gb = [G[f"conv{layer}_{h}.bias"] for h in HEADS]
            contiguous = all(gw[i + 1].data_ptr() - gw[i].data_ptr() == n_out * k_in * 4 for i in range(2))
            if contiguous:
                chk(L.mf_train_gemm_tn(ptr(dz), dz.shape[1], ptr(x), x.shape[1], NP, n_out, k_in,
                                       ptr(gw[0]), k_in, 3, n_out, k_in, n_out * k_in, 1, s()), "gemm_tn")
            else:
                for i in range(3):
                    chk(L.mf_train_gemm_tn(ptr(dz[:, i * n_out:]), dz.shape[1], ptr(x[:, i * k_in:]),
                                           x.shape[1], NP, n_out, k_in, ptr(gw[i]), k_in, 1, 0, 0, 0,
                                           1, s()), "gemm_tn")
            for i in range(3):
                chk(L.mf_train_colsum(ptr(dz[:, i * n_out:]), dz.shape[1], NP, n_out, ptr(gb[i]), s()),
                    "colsum")
            model._gemm_grouped(L, [dict(
                A=dz[:, i * n_out:], W=tw[f"conv{layer}_{h}/Wt"], bias=None, out=dx, M=NP, N=k_in,
                K=n_out, lda=dz.shape[1], ldo=dx.shape[1], col_off=i * k_in, relu=0)
                for i, h in enumerate(HEADS)])
            chk(L.mf_train_relu_mask(ptr(dx), dx.shape[1], ptr(act), act.shape[1], NP, dx.shape[1], s()),
                "relu_mask")
        # ---- layer 1 (the three heads share their input: one [1920 x 984] weight gradient)
        gw1 = [G[f"conv1_{h}.weight"] for h in HEADS]
        if all(gw1[i + 1].data_ptr() - gw1[i].data_ptr() == 640 * 984 * 4 for i in range(2)):
            chk(L.mf_train_gemm_tn(ptr(tb["dhd1"]), 1920, ptr(buf["feat"]), M.FEAT_LD, NP, 1920, 984,
                                   ptr(gw1[0]), 984, 1, 0, 0, 0, 1, s()), "gemm_tn")
        else:
            for i in range(3):
                chk(L.mf_train_gemm_tn(ptr(tb["dhd1"][:, i * 640:]), 1920, ptr(buf["feat"]), M.FEAT_LD,
                                       NP, 640, 984, ptr(gw1[i]), 984, 1, 0, 0, 0, 1, s()), "gemm_tn")
        for i, h in enumerate(HEADS):
            chk(L.mf_train_colsum(ptr(tb["dhd1"][:, i * 640:]), 1920, NP, 640,
                                  ptr(G[f"conv1_{h}.bias"]), s()), "colsum")
        model._gemm(L, tb["dhd1"], tw["head1/Wt"], None, tb["dfeat"], NP, 992, 1920, lda=1920,
                    relu=0, ldo=M.FEAT_LD)
        if on_bucket:
            on_bucket("heads")
        # ---- trilinear gathers backward: scatter into the conv4 / conv3 output grids
        chk(L.mf_train_interp_bwd(ptr(tb["dfeat"]), M.FEAT_LD, 472, ptr(st["points"]), B, P, 512, 8,
                                  4.0, ptr(tb["dgrid4"]), s()), "interp_bwd4")
        chk(L.mf_train_interp_bwd(ptr(tb["dfeat"]), M.FEAT_LD, 216, ptr(st["points"]), B, P, 256, 16,
                                  2.0, ptr(tb["dgrid3"]), s()), "interp_bwd3")
        # ---- conv4: ReLU mask + pad (+ bias gradient), weight gradient, input gradient
        chk(L.mf_train_mask_pack(ptr(tb["dgrid4"]), ptr(buf["h4"]), 0, B, 8, 512, ptr(tb["dY4p"]),
                                 ptr(G["conv4.bias"]), s()), "mask_pack4")
        chk(L.mf_train_conv_wgrad(ptr(tb["dY4p"]), ptr(buf["x4"]), B, 8, 512, 8 * 256, ptr(tb["gw4"]),
                                  0, s()), "conv4 wgrad")
        G["conv4.weight"].add_(unpack_conv_k4s2_grad(tb["gw4"], 512, 256))
        chk(L.mf_train_conv_dgrad(ptr(tb["dY4p"]), ptr(tw["conv4/Wd"]), B, 8, 512, 256, 1,
                                  ptr(tb["dgrid3"]), 256, 0, s()), "conv4 dgrad")
        if on_bucket:
            on_bucket("conv4")
        # ---- conv3
        chk(L.mf_train_mask_pack(ptr(tb["dgrid3"]), ptr(buf["x4"]), 1, B, 16, 256, ptr(tb["dY3p"]),
                                 ptr(G["conv3.bias"]), s()), "mask_pack3")
        chk(L.mf_train_conv_wgrad(ptr(tb["dY3p"]), ptr(buf["x3"]), B, 16, 256, 8 * Ct, ptr(tb["gw3"]),
                                  0, s()), "conv3 wgrad")
        G["conv3.weight"].add_(unpack_conv_k4s2_grad(tb["gw3"], 256, Ct))
        chk(L.mf_train_conv_dgrad(ptr(tb["dY3p"]), ptr(tw["conv3/Wd"]), B, 16, 256, Ct, 2,
                                  ptr(tb["dx3"]), Ct, B * 4096 * Ct, s()), "conv3 dgrad")
        # ---- occupancy stencils (conv1_occ / conv2_occ): 0.24 GFLOP per object, library conv
        if model._with_occupancy:
            _occ_backward(model, st, tb["dx3"], B, D, G)
        # ---- voxelisation backward + per-point MLP backward
        chk(L.mf_train_vox_bwd(ptr(tb["dx3"]), Ct, ptr(buf["prev_keys"]), B, P, 144, D,
                               ptr(tb["dfeat"]), M.FEAT_LD, 72, ptr(tb["dfeat2"]), s()), "vox_bwd")
        dvalues = torch.empty_like(st["values"]) if needs_values else None
        chk(L.mf_train_point_mlp_bwd(
            ptr(st["values"]), ptr(st["points"]), ptr(buf["feat"]), M.FEAT_LD, ptr(tb["dfeat"]),
            M.FEAT_LD, ptr(buf["feat2"]), ptr(tb["dfeat2"]), ptr(tw["conv1_rgb/W"]),
            ptr(tw["conv1_pcd/W"]), ptr(tw["conv2_rgb/W"]), ptr(tw["conv2_pcd/W"]), B, P,
            D / 2.0 - 0.5, ptr(G["conv1_rgb.weight"]), ptr(G["conv1_rgb.bias"]),
            ptr(G["conv1_pcd.weight"]), ptr(G["conv1_pcd.bias"]), ptr(G["conv2_rgb.weight"]),
            ptr(G["conv2_rgb.bias"]), ptr(G["conv2_pcd.weight"]), ptr(G["conv2_pcd.bias"]),
            ptr(dvalues), s()), "point_mlp_bwd")
        if on_bucket:
            on_bucket("conv3+mlp")
    if needs_values:
        G["values"] = dvalues
    return G


def _occ_backward(model, st, dx3, B, D, G):
    """Gradient of conv1_occ / conv2_occ (model.py:114-125) from channels 144..159 of the conv3
    input gradient (parity order -> dense NCDHW), through torch's conv3d."""
    Do = D // 2
    g = dx3[:, :, 144:160].float().reshape(2, 2, 2, B, Do, Do, Do, 16)
    dense = torch.empty((B, D, D, D, 16), dtype=torch.float32, device=dx3.device)
    for rd in range(2):
        for rh in range(2):
            for rw in range(2):
                dense[:, 1 - rd::2, 1 - rh::2, 1 - rw::2] = g[rd, rh, rw]      # x = 2 o + 1 - r
    dense = dense.permute(0, 4, 1, 2, 3)
    F = torch.nn.functional
    ps = [model.conv1_occ.weight, model.conv1_occ.bias, model.conv2_occ.weight, model.conv2_occ.bias]
    with torch.enable_grad():
        leaf = [p.detach().float().requires_grad_(True) for p in ps]
        x = st["gne"].to(torch.float32)[:, None]
        h = F.relu(F.conv3d(x, leaf[0], leaf[1], stride=1, padding=1))
        h = F.relu(F.conv3d(h, leaf[2], leaf[3], stride=1, padding=2, dilation=2))
        gs = torch.autograd.grad(h, leaf, grad_outputs=dense)
    for name, gv in zip(("conv1_occ.weight", "conv1_occ.bias", "conv2_occ.weight", "conv2_occ.bias"), gs):
        G[name].add_(gv)


# ------------------------------------------------------------------ data-parallel trainer
class Trainer:
    """Data-parallel training of the pose model (train.py:229-233,342-344,361): one process per
    GPU, per-rank batch = global batch / world, gradient sum over NCCL, Chainer-form Adam.

    All parameters, gradients and Adam moments are flat fp32 buffers; the model's Parameters and
    their .grad are views.  The all-reduce of a bucket is issued (async, NCCL stream) as soon as
    the backward pass has finished the bucket's gradients and overlaps the rest of the backward;
    after the last bucket one fused kernel per bucket applies 1/world and the Adam update."""

    def __init__(self, model, alpha=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, process_group=None,
                 overlap=True):
        import torch.distributed as dist
        self.model = model
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = process_group
        self.world = self.dist.get_world_size(process_group) if self.dist else 1
        self.overlap = overlap
        self.t = 0
        named = dict(model.named_parameters())
        order = flat_param_order(model)
        rest = [n for n in named if n not in order]              # 2-D extractor: torch autograd
        self.names = order + rest
        sizes = [named[n].numel() for n in self.names]
        offs = np.concatenate([[0], np.cumsum([(s + 3) // 4 * 4 for s in sizes])])   # 16-byte aligned
        total = int(offs[-1])
        dev = named[order[0]].device
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offs = {n: (int(offs[i]), sizes[i]) for i, n in enumerate(self.names)}
        views = {}
        with torch.no_grad():
            for n in self.names:
                o, sz = self.offs[n]
                p = named[n]
                self.flat_p[o:o + sz].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + sz].view(p.shape)
                views[n] = self.flat_g[o:o + sz].view(p.shape)
                p.grad = views[n]
        model._train_grad_views = {n: views[n] for n in order}
        # buckets in backward-completion order
        def span(first, last):
            return self.offs[first][0], self.offs[last][0] + (self.offs[last][1] + 3) // 4 * 4
        b_heads = span(order[0], "conv1_conf.bias")
        b_conv4 = span("conv4.weight", "conv4.bias")
        b_rest3d = span("conv3.weight", order[-1])
        self.buckets = {"heads": b_heads, "conv4": b_conv4, "conv3+mlp": b_rest3d}
        if rest:
            self.buckets["extractor"] = span(rest[0], rest[-1])
        self._handles = []
        model._on_bucket = self._bucket_ready if (self.dist and overlap) else None   # backward hook
        if self.dist:                                   # same initial weights everywhere
            self.dist.broadcast(self.flat_p, src=0, group=self.group)

    def _bucket_ready(self, name):
        lo, hi = self.buckets[name]
        self._handles.append((name, self.dist.all_reduce(self.flat_g[lo:hi], group=self.group,
                                                         async_op=True)))

    def alpha_t(self):
        import math
        return self.alpha * math.sqrt(1.0 - math.pow(self.beta2, self.t)) / \
            (1.0 - math.pow(self.beta1, self.t))

    def zero_grad(self):
        self.flat_g.zero_()

    def reduce_gradients(self):
        """All-reduce (sum) every bucket that the backward pass has not already issued, then wait
        for all of them.  No-op on a single rank."""
        if not self.dist:
            return
        done = {n for n, _ in self._handles}
        for name, (lo, hi) in self.buckets.items():
            if name not in done:
                self._handles.append((name, self.dist.all_reduce(
                    self.flat_g[lo:hi], group=self.group, async_op=True)))
        for _, h in self._handles:
            h.wait()
        self._handles = []

    def update(self):
        """Fused 1/world unscale + Chainer-form Adam over the flat buffers (one launch)."""
        L = _lib.lib()
        self.t += 1
        n = self.flat_p.numel()
        with torch.cuda.device(self.flat_p.device):
            _lib.check(L.mf_train_adam(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.flat_m),
                                       _lib.ptr(self.flat_v), n, float(np.float32(self.alpha_t())),
                                       self.beta1, self.beta2, self.eps, 1.0, 1.0 / self.world,
                                       _lib.stream()), "adam")
        # the update ran outside torch: the kernels' packed bf16 copies must be rebuilt
        self.model._packed_ver = None
        self.model._train_packed = None

    def step(self, **batch):
        """One training step: forward, loss, backward (gradients into the flat buffer), bucketed
        all-reduce, fused unscale + Adam.  Returns the (local) loss tensor."""
        self.zero_grad()
        self._handles = []
        loss = self.model(**batch)
        loss.backward()
        self.reduce_gradients()
        self.update()
        return loss
