"""singleview_3d pose model: the 3D-CNN section on hand-written sm_100a kernels.

Mirrors morefusion/contrib/singleview_3d/models/model.py (class Model :11-481): same constructor
keywords (:19-27), the same link names for every layer (:62-91) so a reference snapshot maps 1:1
onto ``state_dict`` keys, and the same public methods:
  * ``predict(class_id, rgb, pcd, pitch, origin, grid_nontarget_empty)`` (:166-275): NaN mask,
    2-D extractor (torch / cuDNN: the adjacent row SURVEY.md 8f-1), per-object 1000-point
    sampling with the reference's RandomState(1234) permutation, default pitch / origin,
    voxel-frame transform, then ``forward_features`` -- the hot path proper: everything after
    the 2-D extractor on this library's kernels;
  * ``evaluate`` (:325-375), ``loss`` (:377-481, without the "+occupancy" terms the reference
    itself calls with a stale signature) and ``__call__`` (:277-323).

Internal layout is B200-first, not a translation: activations are channels-last bf16 in
persistent buffers, the two k4/s2 Conv3Ds and all Conv1D heads are GEMMs on one kernel
family (tcgen05 when the shape qualifies, SIMT otherwise), and the whole forward is a fixed
launch sequence that can be captured in a CUDA graph.  No cuDNN / cuBLAS / torch ops on the
path after the 2-D feature extractor.
"""

import ctypes

import numpy as np
import torch

from .... import _lib
from ....functions.geometry import _util
from ....functions.geometry.average_voxelization_3d import AverageVoxelization3D

# YCB voxel pitch for a 32^3 grid = bbox diagonal / 32
# (datasets/ycb_video/models.py:113-115; table ros/.../utils/data.h:12-32), index = class id
YCB_VOXEL_PITCH_32 = [
    None, 0.006296589104319322, 0.008705823111730123, 0.006425726070431774,
    0.004375644727606043, 0.007023497839423789, 0.003923674166124662, 0.006018916012848706,
    0.004320481778555272, 0.004535342826373148, 0.006631487204390293, 0.009982031658204186,
    0.008721623259758258, 0.007331656585392745, 0.005318687227615036, 0.008406278399464109,
    0.0079006960844688, 0.00699458097945295, 0.0038783371057780278, 0.006648125743278138,
    0.008405508709996566, 0.0033429720217908734,
]


class GemmParams(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("W", ctypes.c_void_p), ("bias", ctypes.c_void_p),
        ("out", ctypes.c_void_p), ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("mode", ctypes.c_int), ("lda", ctypes.c_longlong), ("ldw", ctypes.c_longlong),
        ("Do", ctypes.c_int), ("Ci8", ctypes.c_int), ("relu", ctypes.c_int),
        ("out_mode", ctypes.c_int), ("ldo", ctypes.c_longlong), ("col_off", ctypes.c_int),
    ]


GEMM_LINEAR, GEMM_CONV_S2D = 0, 1
FEAT_LD = 1024          # row pitch (elements) of the [points, 984] concat feature and head-1 weights
OUT_BF16, OUT_F32, OUT_S2D_BF16 = 0, 1, 2


def pack_conv_k4s2_weight(W):
    """OIDHW [Co,Ci,4,4,4] fp32 -> [Co, 64*Ci] bf16 in implicit-GEMM K order
    k = ((ad*2+ah)*2+aw)*8Ci + (rd*4+rh*2+rw)*Ci + ci with kd = 2ad+rd (etc.)."""
    Co, Ci = W.shape[:2]
    W = W.reshape(Co, Ci, 2, 2, 2, 2, 2, 2)            # co ci ad rd ah rh aw rw
    W = W.permute(0, 2, 4, 6, 3, 5, 7, 1)              # co ad ah aw rd rh rw ci
    return W.reshape(Co, 64 * Ci).to(torch.bfloat16).contiguous()


class Model(torch.nn.Module):

    _lambda_confidence = 0.015
    _n_point = 1000
    _voxel_dim = 32

    def __init__(self, *, n_fg_class, pretrained_resnet18=False, with_occupancy=False,
                 loss=None, loss_scale=None):
        super().__init__()
        self._n_fg_class = n_fg_class
        self._with_occupancy = with_occupancy
        if loss is None:
            loss = "add/add_s"
        assert loss in ["add", "add/add_s", "add+occupancy", "add/add_s+occupancy"]
        self._loss = loss
        if loss_scale is None:
            loss_scale = {"occupancy": 1.0}
        self._loss_scale = loss_scale
        self._pretrained_resnet18 = pretrained_resnet18

        nn = torch.nn
        self.conv1_rgb = nn.Conv1d(32, 64, 1)
        self.conv1_pcd = nn.Conv1d(3, 8, 1)
        self.conv2_rgb = nn.Conv1d(64, 128, 1)
        self.conv2_pcd = nn.Conv1d(8, 16, 1)
        if with_occupancy:
            self.conv1_occ = nn.Conv3d(1, 8, 3, 1, padding=1)
            self.conv2_occ = nn.Conv3d(8, 16, 3, 1, padding=2, dilation=2)
        cin3 = 144 + (16 if with_occupancy else 0)
        self.conv3 = nn.Conv3d(cin3, 256, 4, 2, padding=1)
        self.conv4 = nn.Conv3d(256, 512, 4, 2, padding=1)
        for head, cout in (("rot", n_fg_class * 4), ("trans", n_fg_class * 3),
                           ("conf", n_fg_class)):
            setattr(self, f"conv1_{head}", nn.Conv1d(984, 640, 1))
            setattr(self, f"conv2_{head}", nn.Conv1d(640, 256, 1))
            setattr(self, f"conv3_{head}", nn.Conv1d(256, 128, 1))
            setattr(self, f"conv4_{head}", nn.Conv1d(128, cout, 1))
        # 2-D feature extractor (model.py:42-60) and the CAD model source (:29)
        from ....models import ResNet18Extractor
        from ....models.dense_fusion import PSPNetExtractor
        from .... import synthetic
        self.resnet_extractor = ResNet18Extractor()
        self.pspnet_extractor = PSPNetExtractor()
        self._models = synthetic.SyntheticYCBModels()
        self.reported = {}
        self._pending_eval = None
        self._raw8 = None
        self._packed = None
        self._packed_ver = None
        self._wbufs = {}
        self.use_tensor_cores = True
        # "bf16": throughput mode (bf16 operands / stored activations, fp32 accumulation);
        # "bf16x3": fp32-class parity mode -- every GEMM operand split hi + lo and each product
        # evaluated as A_hi W_hi + A_lo W_hi + A_hi W_lo on the same tcgen05 kernels
        self.precision = "bf16"
        self.fused_voxelize = True
        # independent branches (occupancy stencil || point MLP + voxelisation; conv3-level gather
        # || conv4) run on a second stream; captured into the CUDA graphs as parallel branches
        self.fused_head4 = True     # last head layer + class select + pose epilogue in one kernel
        # inference: up3 + 1x1 conv + log-softmax of the 2-D extractor only at the sampled pixels
        # (csrc/extractor_tail.cu); training keeps the dense torch path (dropout, autograd)
        self.fused_extractor_tail = True
        self.concurrent_branches = True
        self.fused_occ = True       # conv1_occ + conv2_occ in one kernel (no global intermediate)
        self.stream_k = True        # conv3 / conv4: equal K-block ranges per SM, reduction in the epilogue
        self.fused_heads = True     # head layers 1-3 (7 GEMMs) as one persistent launch with tile-level dependencies
        self._side_streams = {}
        self.launch_log = []
        self.n_launches = 0      # kernels of this library launched so far (bench's gpu_launches)

    # ------------------------------------------------------------------ weights
    def load_reference_weights(self, weights):
        """weights: {'<link>/W': ndarray, '<link>/b': ndarray} keyed like a chainer npz
        snapshot of the reference model (model.py:62-91)."""
        with torch.no_grad():
            for name, mod in self.named_children():
                if name + "/W" in weights:
                    mod.weight.copy_(torch.as_tensor(weights[name + "/W"]).reshape(mod.weight.shape))
                    mod.bias.copy_(torch.as_tensor(weights[name + "/b"]))
        self._packed_ver = None
        return self

    def _pack(self):
        dev = self.conv3.weight.device
        f32 = lambda t: t.detach().to(torch.float32).contiguous()      # noqa: E731
        p = {}
        for n in ("conv1_rgb", "conv1_pcd", "conv2_rgb", "conv2_pcd"):
            m = getattr(self, n)
            # k-major ([in][out]) for k_point_mlp's shared-memory staging
            p[n + "/W"] = f32(m.weight.reshape(m.weight.shape[0], -1).t())
            p[n + "/b"] = f32(m.bias)
        if self._with_occupancy:
            for n in ("conv1_occ", "conv2_occ"):
                m = getattr(self, n)
                p[n + "/W"] = f32(m.weight)
                p[n + "/b"] = f32(m.bias)
        for n in ("conv3", "conv4"):
            m = getattr(self, n)
            p[n + "/W"] = pack_conv_k4s2_weight(m.weight.detach().float())
            p[n + "/b"] = f32(m.bias)
        # heads: layer 1 of the three heads share their input -> one GEMM with N = 3*640
        heads = ("rot", "trans", "conf")
        # row pitch padded 984 -> FEAT_LD so every 128-byte TMA row segment is line-aligned
        # (K stays 984 in the tensor maps; the pad columns are never read)
        W1 = torch.cat(
            [getattr(self, f"conv1_{h}").weight.detach().reshape(640, 984) for h in heads], 0)
        p["head1/W"] = torch.nn.functional.pad(W1, (0, FEAT_LD - 984)).to(torch.bfloat16).contiguous()
        p["head1/b"] = torch.cat([f32(getattr(self, f"conv1_{h}").bias) for h in heads])
        for h in heads:
            for layer in (2, 3, 4):
                m = getattr(self, f"conv{layer}_{h}")
                W = m.weight.detach().reshape(m.weight.shape[0], -1).to(torch.bfloat16)
                if layer == 4:     # pad N to a multiple of 8 rows? not needed: K-major rows only
                    pass
                p[f"conv{layer}_{h}/W"] = W.contiguous()
                p[f"conv{layer}_{h}/b"] = f32(m.bias)
        old = self._packed
        if (old is not None and self._packed_dev == dev and old.keys() == p.keys()
                and all(old[k].shape == p[k].shape and old[k].dtype == p[k].dtype for k in p)):
            # refresh in place: captured CUDA graphs keep reading the same addresses
            for k in p:
                old[k].copy_(p[k])
            return old
        self._packed = p
        self._packed_dev = dev
        return p

    def _work_buffers(self, B, P, dev):
        key = (B, P, dev)
        if key in self._wbufs:
            return self._wbufs[key]
        D = self._voxel_dim
        Ct = 144 + (16 if self._with_occupancy else 0)
        bf, f32 = torch.bfloat16, torch.float32
        z = lambda *s, dt=bf: torch.zeros(*s, dtype=dt, device=dev)     # noqa: E731
        NP = B * P
        b = dict(
            feat=z(NP, FEAT_LD), feat2=z(NP, 144, dt=f32),
            x3=z(B, 17, 17, 17, 8 * Ct),          # s2d of the zero-padded 32^3 x Ct grid
            x4=z(B, 9, 9, 9, 8 * 256),            # s2d of the zero-padded 16^3 x 256 grid (= H3)
            h4=z(B, 8, 8, 8, 512),
            hd1=z(NP, 1920), hd2=z(NP, 3 * 256), hd3=z(NP, 3 * 128),
            out_rot=z(NP, self._n_fg_class * 4, dt=f32),
            out_trans=z(NP, self._n_fg_class * 3, dt=f32),
            out_conf=z(NP, self._n_fg_class, dt=f32),
            bi=torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(P),
            prev_keys=torch.full((2 * NP,), -1, dtype=torch.int32, device=dev),   # keys | sorted order
            # stream-K GEMMs (conv3 / conv4): flag words (zeroed once, the kernels keep them
            # zero) and the per-SM fp32 accumulator slots, owned by this buffer set
            sk_sync=torch.zeros(2048, dtype=torch.int32, device=dev),
            hd_sync=torch.zeros(4096, dtype=torch.int32, device=dev),      # fused heads: epoch + counters
            sk_ws=torch.empty(256 * 128 * 256, dtype=f32, device=dev),
        )
        if self._with_occupancy:
            b["occ1"] = z(B, D ** 3, 8, dt=f32)
            b["occ1_bf16"] = z(B, D ** 3, 8)
            b["occ2"] = z(B, D ** 3, 16, dt=f32)
        self._wbufs[key] = b
        return b

    # ------------------------------------------------------------------ kernels
    def _gemm(self, L, A, W, bias, out, M, N, K, *, mode=GEMM_LINEAR, lda=0, Do=0, Ci8=0,
              relu=1, out_mode=OUT_BF16, ldo=0, col_off=0, streamk=None):
        gp = GemmParams(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(out), M, N, K, mode,
                        lda, W.shape[1], Do, Ci8, relu, out_mode, ldo, col_off)
        if self.use_tensor_cores:
            if streamk is not None and self.stream_k:
                sync, ws = streamk
                rc = L.mf_gemm_bf16_tc_ex(ctypes.byref(gp), 1, _lib.ptr(ws), ws.numel() * ws.element_size(),
                                          _lib.ptr(sync), None, 0, _lib.stream())
                launches = 1
            else:
                ws = _util.workspace(L.mf_gemm_bf16_tc_workspace_bytes(M, N), A.device)
                rc = L.mf_gemm_bf16_tc(ctypes.byref(gp), _lib.ptr(ws), ws.numel(), _lib.stream())
                launches = 1 + (1 if (M // 128) * max(N // 256, 1) <= 74 and K >= 4096 else 0)
            if rc == 0:
                self.launch_log.append(("tc", M, N, K))
                self.n_launches += launches
                return
            if rc != -4:          # MF_E_UNSUPPORTED -> the SIMT kernel covers the shape
                _lib.check(rc, "gemm_bf16_tc")
        self.launch_log.append(("simt", M, N, K))
        self.n_launches += 1
        _lib.check(L.mf_gemm_bf16_simt(ctypes.byref(gp), _lib.stream()), "gemm_bf16_simt")

    def _gemm_grouped(self, L, specs, streamk=None):
        """Up to 3 GEMMs in one launch; tcgen05 when the shapes match and qualify, else SIMT."""
        n = len(specs)
        arr = (GemmParams * n)()
        keep = []
        for i, sp in enumerate(specs):
            keep.append(sp)
            arr[i] = GemmParams(
                _lib.ptr(sp["A"]), _lib.ptr(sp["W"]), _lib.ptr(sp["bias"]), _lib.ptr(sp["out"]),
                sp["M"], sp["N"], sp["K"], sp.get("mode", GEMM_LINEAR), sp.get("lda", 0),
                sp["W"].shape[1], sp.get("Do", 0), sp.get("Ci8", 0), sp.get("relu", 1),
                sp.get("out_mode", OUT_BF16), sp.get("ldo", 0), sp.get("col_off", 0))
        if self.use_tensor_cores:
            dev = specs[0]["A"].device
            if streamk is not None and self.stream_k:
                sync, ws = streamk
                rc = L.mf_gemm_bf16_tc_ex(arr, n, _lib.ptr(ws), ws.numel() * ws.element_size(),
                                          _lib.ptr(sync), None, 0, _lib.stream())
            else:
                ws = _util.workspace(L.mf_gemm_bf16_tc_workspace_bytes(specs[0]["M"], specs[0]["N"]), dev)
                rc = L.mf_gemm_bf16_tc_grouped(arr, n, _lib.ptr(ws), ws.numel(), _lib.stream())
            if rc == 0:
                self.n_launches += 1
                return
            if rc != -4:
                _lib.check(rc, "gemm_bf16_tc_grouped")
        _lib.check(L.mf_gemm_bf16_simt_grouped(arr, n, _lib.stream()), "gemm_bf16_simt_grouped")
        self.n_launches += 1

    def _heads_fused(self, L, buf, w, NP):
        """Layers 1-3 of the three heads as one persistent launch (mf_cnn_heads_tc).  Returns
        False when the library declines the shapes (the caller then runs the three launches)."""
        heads = ("rot", "trans", "conf")
        arr = (GemmParams * 7)()
        keep = [buf["hd1"][:, i * 640:] for i in range(3)] + [buf["hd2"][:, i * 256:] for i in range(3)]
        arr[0] = GemmParams(_lib.ptr(buf["feat"]), _lib.ptr(w["head1/W"]), _lib.ptr(w["head1/b"]),
                            _lib.ptr(buf["hd1"]), NP, 1920, 984, GEMM_LINEAR, FEAT_LD,
                            w["head1/W"].shape[1], 0, 0, 1, OUT_BF16, 1920, 0)
        for i, h in enumerate(heads):
            W2, W3 = w[f"conv2_{h}/W"], w[f"conv3_{h}/W"]
            arr[1 + i] = GemmParams(_lib.ptr(keep[i]), _lib.ptr(W2), _lib.ptr(w[f"conv2_{h}/b"]),
                                    _lib.ptr(buf["hd2"]), NP, 256, 640, GEMM_LINEAR, 1920, W2.shape[1],
                                    0, 0, 1, OUT_BF16, 768, i * 256)
            arr[4 + i] = GemmParams(_lib.ptr(keep[3 + i]), _lib.ptr(W3), _lib.ptr(w[f"conv3_{h}/b"]),
                                    _lib.ptr(buf["hd3"]), NP, 128, 256, GEMM_LINEAR, 768, W3.shape[1],
                                    0, 0, 1, OUT_BF16, 384, i * 128)
        rc = L.mf_cnn_heads_tc(arr, 7, _lib.ptr(buf["hd_sync"]), _lib.stream())
        if rc == -4:
            return False
        _lib.check(rc, "heads_tc")
        self.launch_log.append(("tc", NP, 1920, 984))
        self.n_launches += 1
        return True

    # ------------------------------------------------------------------ reference API
    def _keep_indices(self, n_point):
        """model.py:206-219: which of the n valid pixels feed the network."""
        n_point = int(n_point)
        if n_point <= 0:
            raise ValueError("an object has no valid (non-NaN) point")
        rs = np.random.mtrand._rand if self.training else np.random.RandomState(1234)
        if n_point >= self._n_point:
            return rs.permutation(n_point)[:self._n_point]
        return np.r_[np.arange(n_point), rs.randint(0, n_point, self._n_point - n_point)]

    def predict(self, *, class_id, rgb, pcd, pitch=None, origin=None, grid_nontarget_empty=None):
        """rgb [B,H,W,3] uint8, pcd [B,H,W,3] f32 with NaN at invalid pixels, class_id [B],
        pitch [B] / origin [B,3] (entries or the whole argument may be None -> YCB voxel pitch,
        median of the object's points - 15.5 pitch), grid_nontarget_empty [B,32,32,32].
        Returns rot [B,P,4], trans [B,P,3], conf [B,P] (model.py:166-275).

        One device->host read per call (the valid-pixel counts of the batch, which size the
        reference's host-side permutation); the reference syncs several times per object."""
        dev = self.conv3.weight.device
        if dev.type != "cuda":
            raise RuntimeError("Model runs on CUDA only (no CPU fallback); call .cuda() first")
        f32 = torch.float32
        rgb = torch.as_tensor(rgb, device=dev)
        pcd = torch.as_tensor(pcd, device=dev).to(f32)
        B, H, W, _ = rgb.shape
        P = self._n_point
        class_id_h = None if isinstance(class_id, torch.Tensor) and class_id.is_cuda \
            else np.asarray(class_id).reshape(B)
        if class_id_h is not None and ((class_id_h < 1) | (class_id_h > self._n_fg_class)).any():
            raise IndexError("class_id out of range for n_fg_class")     # fancy index at :266-269
        mask = ~torch.isnan(pcd).any(dim=3)                              # [B,H,W]   (:178)
        sparse_tail = (self.fused_extractor_tail and not self.training
                       and not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())))
        h_res = self.resnet_extractor(rgb.permute(0, 3, 1, 2).to(f32))
        if sparse_tail:
            h_up2 = self.pspnet_extractor.forward_up2(h_res)             # [B,64,H/2,W/2]
            if h_up2.shape[2] * 2 != H or h_up2.shape[3] * 2 != W:
                raise ValueError("image sides must be multiples of 8 (the extractor returns H x W features)")
        else:
            h_rgb = self.pspnet_extractor.up3(self.pspnet_extractor.forward_up2(h_res))
            h_rgb = torch.log_softmax(self.pspnet_extractor.conv1(h_rgb), dim=1)
        flat_mask = mask.flatten(1)
        n_point = flat_mask.sum(1).cpu().numpy()                         # the one host read
        keep = torch.as_tensor(np.stack([self._keep_indices(n) for n in n_point]), device=dev)
        cums = flat_mask.to(torch.int64).cumsum(1)
        pix = torch.searchsorted(cums, keep + 1)                         # k-th valid pixel, row-major
        pcd_f = pcd.permute(0, 3, 1, 2).flatten(2)                       # [B,3,HW]
        points = pcd_f.gather(2, pix[:, None, :].expand(B, 3, P))
        if sparse_tail:
            values = self._extractor_tail(h_up2, pix)
        else:
            values = h_rgb.flatten(2).gather(2, pix[:, None, :].expand(B, h_rgb.shape[1], P))
        # defaults (:197-205)
        pitch_l = [None] * B if pitch is None else list(pitch)
        if any(p is None for p in pitch_l):
            if class_id_h is None:
                class_id_h = class_id.cpu().numpy()
            pitch_l = [self._models.get_voxel_pitch(self._voxel_dim, class_id_h[i])
                       if pitch_l[i] is None else float(pitch_l[i]) for i in range(B)]
            pitch_t = torch.tensor(pitch_l, dtype=f32, device=dev)
        else:
            pitch_t = torch.as_tensor(pitch, device=dev).to(f32).reshape(B)
        origin_l = [None] * B if origin is None else list(origin)
        if any(o is None for o in origin_l):
            # extra/_cupy.py:47-62 median over the object's valid points, per axis
            srt = torch.where(flat_mask[:, None, :], pcd_f, torch.full_like(pcd_f, float("inf")))
            srt = srt.sort(dim=2).values
            n = torch.as_tensor(n_point, device=dev)
            hi = srt.gather(2, (n // 2)[:, None, None].expand(B, 3, 1))[:, :, 0]
            lo = srt.gather(2, ((n - 1) // 2)[:, None, None].expand(B, 3, 1))[:, :, 0]
            center = torch.where((n % 2 == 1)[:, None], hi, (hi + lo) / 2)
            auto = center - pitch_t[:, None] * (self._voxel_dim / 2.0 - 0.5)
            given = torch.stack([auto[i] if origin_l[i] is None
                                 else torch.as_tensor(origin_l[i], device=dev).to(f32)
                                 for i in range(B)])
            origin_t = given
        else:
            origin_t = torch.as_tensor(origin, device=dev).to(f32).reshape(B, 3)
        points = (points - origin_t[:, :, None]) / pitch_t[:, None, None]      # (:236)
        return self._features(class_id=class_id, values=values, points=points, pitch=pitch_t,
                              origin=origin_t, grid_nontarget_empty=grid_nontarget_empty)

    def _extractor_tail(self, h_up2, pix):
        """up3 + conv1 + log-softmax of the PSPNet upsampler (pspnet.py:64-82) at the sampled pixels
        only: h_up2 [B,64,Hs,Ws], pix [B,P] int64 -> values [B,32,P] (csrc/extractor_tail.cu)."""
        ex = self.pspnet_extractor
        ps = (ex.up3.conv.weight, ex.up3.conv.bias, ex.up3.prelu.weight, ex.conv1.weight, ex.conv1.bias)
        ver = tuple((p.data_ptr(), p._version) for p in ps)
        cache = self.__dict__.get("_tail_pack")
        if cache is None or cache[0] != ver:
            with torch.no_grad():
                w3t = ps[0].detach().float().reshape(64, 576).t().contiguous()
                w1t = ps[3].detach().float().reshape(32, 64).t().contiguous()
                packed = (w3t, ps[1].detach().float().contiguous(), ps[2].detach().float().contiguous(),
                          w1t, ps[4].detach().float().contiguous())
            cache = (ver, packed)
            self.__dict__["_tail_pack"] = cache
        w3t, b3, slope, w1t, b1 = cache[1]
        B, C, Hs, Ws = h_up2.shape
        assert C == 64 and slope.numel() == 1
        x = h_up2.detach().float().permute(0, 2, 3, 1).contiguous()      # channels-last rows
        pix = pix.to(torch.int64).contiguous()
        P = pix.shape[1]
        out = torch.empty((B, 32, P), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mf_psp_tail_sampled(
                _lib.ptr(x), _lib.ptr(pix), B, P, Hs, Ws, _lib.ptr(w3t), _lib.ptr(b3), _lib.ptr(slope),
                _lib.ptr(w1t), _lib.ptr(b1), _lib.ptr(out), _lib.stream()), "psp_tail_sampled")
        return out

    def _features(self, **kw):
        """forward_features with autograd when gradients are enabled (training)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from . import training
            return training.forward_features_with_grad(self, **kw)
        return self.forward_features(**kw)

    def evaluate(self, *, class_id, quaternion_true, translation_true, quaternion_pred,
                 translation_pred):
        """ADD / ADD-S of the predicted poses (model.py:325-375); the summary chainer would
        ``report`` is kept in ``self.reported`` and returned.  In training mode the metric
        kernels run on a side stream and the summary is materialised by ``flush_reports()`` (or
        the next ``evaluate``): no device->host read inside the training step."""
        from .... import functions, metrics
        from ....metrics.average_distance import average_distance_device
        dev = self.conv3.weight.device
        t = lambda x, tag: _util.h2d(x, dev, torch.float32, tag).detach()   # noqa: E731
        cid = np.asarray(class_id.cpu() if isinstance(class_id, torch.Tensor) else class_id)
        B = cid.shape[0]
        self.flush_reports()
        main = torch.cuda.current_stream(dev)
        side = self._side(dev, 2) if self.training else main
        q_t, t_t = t(quaternion_true, "eval/q"), t(translation_true, "eval/t")
        q_p, t_p = t(quaternion_pred, "eval/qp"), t(translation_pred, "eval/tp")
        side.wait_stream(main)
        with torch.cuda.stream(side):
            T_true = functions.transformation_matrix(q_t, t_t)
            T_pred = functions.transformation_matrix(q_p, t_p)
            res = average_distance_device(
                [self._pcd_device(int(c), dev) for c in cid],
                [T_true[i] for i in range(B)], [T_pred[i] for i in range(B)], device=dev)
            done = torch.cuda.Event()
            done.record(side)
        for x in (q_t, t_t, q_p, t_p):
            x.record_stream(side)
        self._pending_eval = (res, cid, done, self.training)
        if self.training:
            return None
        return self.flush_reports()

    def _pcd_device(self, class_id, dev):
        """CAD point cloud of a class, cached on the device (uploaded once)."""
        key = (class_id, dev)
        cache = self.__dict__.setdefault("_pcd_cache", {})
        if key not in cache:
            cache[key] = torch.as_tensor(np.ascontiguousarray(
                self._models.get_pcd(class_id=class_id), dtype=np.float32), device=dev)
        return cache[key]

    def flush_reports(self):
        """Materialise a pending ``evaluate`` summary into ``self.reported`` (one host read)."""
        pend = getattr(self, "_pending_eval", None)
        if pend is None:
            return None
        res, cid, done, training = pend
        self._pending_eval = None
        done.synchronize()
        r = res.cpu().numpy().astype(float)
        adds, add_ss = r[:, 0], r[:, 1]
        sym = np.isin(cid, self._models.class_ids_symmetric)
        add_or = np.where(sym, add_ss, adds)
        if training:
            summary = {"add": float(adds.mean()), "add_s": float(add_ss.mean()),
                       "add_or_add_s": float(add_or.mean())}
        else:
            summary = {}
            for i in range(len(cid)):
                key = f"{int(cid[i]):04d}/{i}"
                summary[f"add/{key}"] = float(adds[i])
                summary[f"add_s/{key}"] = float(add_ss[i])
                summary[f"add_or_add_s/{key}"] = float(add_or[i])
        self.reported.update(summary)
        return summary

    def loss(self, class_id, quaternion_true, translation_true, quaternion_pred,
             translation_pred, confidence_pred, pitch=None, origin=None, grid_target=None,
             grid_nontarget_empty=None):
        """Confidence-weighted ADD / ADD-S loss (model.py:377-441, :475-481): per object
        mean over the points with conf > 0 of add * conf - lambda log(conf), averaged over the
        batch.  No boolean indexing (no host read): masked sums."""
        from .... import functions
        if self._loss in ("add+occupancy", "add/add_s+occupancy"):
            raise NotImplementedError(
                "the '+occupancy' loss terms call pseudo_occupancy_voxelization with a stale "
                "signature in the reference (model.py:454-459) and cannot run there either")
        dev = quaternion_pred.device
        quaternion_true = _util.h2d(quaternion_true, dev, torch.float32, "loss/q")
        translation_true = _util.h2d(translation_true, dev, torch.float32, "loss/t")
        cid = np.asarray(class_id.cpu() if isinstance(class_id, torch.Tensor) else class_id)
        B, P = confidence_pred.shape
        T_pred = functions.transformation_matrix(
            quaternion_pred.reshape(B * P, 4), translation_pred.reshape(B * P, 3)).reshape(B, P, 4, 4)
        T_true = functions.transformation_matrix(quaternion_true, translation_true)     # [B,4,4]
        # 500 random CAD points per object (model.py:416-418): indices drawn on the host with
        # numpy's global RNG as in the reference, one upload for the batch, gather on the device
        pcds = [self._pcd_device(int(cid[i]), dev) for i in range(B)]
        sel = _util.h2d(np.stack([np.random.permutation(p.shape[0])[:500] for p in pcds]), dev,
                        torch.int64, "loss/sel")
        cads = torch.stack([pcds[i][sel[i]] for i in range(B)])
        syms = []
        for i in range(B):
            sym = int(cid[i]) in self._models.class_ids_symmetric
            if self._loss == "add":
                sym = False
            elif self._loss == "add_s":
                sym = True
            syms.append(sym)
        from ....functions.loss.average_distance import average_distance_batched
        add = average_distance_batched(cads, T_true, T_pred, syms)           # [B,P]
        conf = confidence_pred
        keep = conf.detach() > 0
        safe = torch.where(keep, conf, torch.ones_like(conf))
        term = torch.where(keep, add * conf - self._lambda_confidence * torch.log(safe),
                           torch.zeros_like(conf))
        loss_i = term.sum(dim=1) / keep.sum(dim=1).clamp(min=1)
        loss = loss_i.mean()
        self.reported["loss"] = loss.detach()      # a tensor: float() it when reporting
        return loss

    def forward(self, *, class_id, rgb, pcd, quaternion_true, translation_true, pitch=None,
                origin=None, grid_target=None, grid_nontarget_empty=None):
        """Training call (model.py:277-323): predict -> evaluate (argmax-confidence pose) -> loss."""
        rot, trans, conf = self.predict(class_id=class_id, rgb=rgb, pcd=pcd, pitch=pitch,
                                        origin=origin, grid_nontarget_empty=grid_nontarget_empty)
        B = rot.shape[0]
        idx = conf.detach().argmax(dim=1)
        ar = torch.arange(B, device=rot.device)
        with torch.no_grad():
            self.evaluate(class_id=class_id, quaternion_true=quaternion_true,
                          translation_true=translation_true,
                          quaternion_pred=rot.detach()[ar, idx],
                          translation_pred=trans.detach()[ar, idx])
        return self.loss(class_id=class_id, quaternion_true=quaternion_true,
                         translation_true=translation_true, quaternion_pred=rot,
                         translation_pred=trans, confidence_pred=conf, pitch=pitch,
                         origin=origin, grid_target=grid_target,
                         grid_nontarget_empty=grid_nontarget_empty)

    def forward_features(self, *, class_id, values, points, pitch, origin,
                         grid_nontarget_empty=None):
        """The hot path proper: everything after the 2-D extractor (model.py:236-273).

        values [B,32,P] f32 per-point RGB features, points [B,3,P] f32 in the voxel frame,
        pitch [B], origin [B,3], class_id [B] (int32), grid_nontarget_empty [B,32,32,32].
        Returns rot [B,P,4], trans [B,P,3], conf [B,P] (fp32)."""
        _lib.require_cuda(values, points)
        dev = values.device
        st = dict(
            values=values.contiguous().float(), points=points.contiguous().float(),
            class_id=_util.h2d(class_id, dev, torch.int32, "ff/class_id").contiguous(),
            pitch=torch.as_tensor(pitch, dtype=torch.float32, device=dev).contiguous(),
            origin=torch.as_tensor(origin, dtype=torch.float32, device=dev).contiguous(),
            gne=None if grid_nontarget_empty is None
            else torch.as_tensor(grid_nontarget_empty, device=dev))
        B, _, P = values.shape
        out = dict(rot=torch.empty((B, P, 4), dtype=torch.float32, device=dev),
                   trans=torch.empty((B, P, 3), dtype=torch.float32, device=dev),
                   conf=torch.empty((B, P), dtype=torch.float32, device=dev))
        if self.precision == "bf16x3":
            self._forward_precise(st, out)
            return out["rot"], out["trans"], out["conf"]
        assert self.precision == "bf16", self.precision
        self._stage_pre(st)
        self._stage_conv3(st)
        self._stage_post(st, out)
        return out["rot"], out["trans"], out["conf"]

    # ------------------------------------------------------------------ fp32-class parity mode
    def _pack_precise(self):
        ver = tuple(p._version for p in self.parameters())
        pk = getattr(self, "_packed_px", None)
        if pk is not None and pk["ver"] == ver:
            return pk
        bf, f32 = torch.bfloat16, torch.float32

        def hilo(W):
            hi = W.to(bf)
            return hi.contiguous(), (W - hi.to(f32)).to(bf).contiguous()
        p = dict(ver=ver)
        for n in ("conv3", "conv4"):
            W = getattr(self, n).weight.detach().float()
            Co, Ci = W.shape[:2]
            Wp = W.reshape(Co, Ci, 2, 2, 2, 2, 2, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(Co, 64 * Ci)
            p[n + "/hi"], p[n + "/lo"] = hilo(Wp)
        heads = ("rot", "trans", "conf")
        W1 = torch.cat([getattr(self, f"conv1_{h}").weight.detach().float().reshape(640, 984)
                        for h in heads], 0)
        p["head1/hi"], p["head1/lo"] = hilo(torch.nn.functional.pad(W1, (0, FEAT_LD - 984)))
        for h in heads:
            for layer in (2, 3):
                m = getattr(self, f"conv{layer}_{h}")
                p[f"conv{layer}_{h}/hi"], p[f"conv{layer}_{h}/lo"] = hilo(
                    m.weight.detach().float().reshape(m.weight.shape[0], -1))
            m = getattr(self, f"conv4_{h}")
            p[f"conv4_{h}/W32"] = m.weight.detach().float().reshape(m.weight.shape[0], -1).contiguous()
        self._packed_px = p
        return p

    def _forward_precise(self, st, out):
        """model.py:93-141,:239-273 at fp32-class accuracy (see csrc/precise.cu)."""
        from .... import config
        L, dev, B, P, w, _ = self._ctx(st)
        px = self._pack_precise()
        D, NP, nfg = self._voxel_dim, B * P, self._n_fg_class
        Cocc = 16 if self._with_occupancy else 0
        Ct = 144 + Cocc
        s, ptr, chk = _lib.stream, _lib.ptr, _lib.check
        key = ("px", B, P, dev)
        if key not in self._wbufs:
            bf, f32 = torch.bfloat16, torch.float32
            z = lambda *sh, dt=bf: torch.zeros(*sh, dtype=dt, device=dev)     # noqa: E731
            self._wbufs[key] = dict(
                feat=[z(NP, FEAT_LD), z(NP, FEAT_LD)], feat1=z(NP, 72, dt=f32), feat2=z(NP, 144, dt=f32),
                x3=[z(B, 17, 17, 17, 8 * Ct), z(B, 17, 17, 17, 8 * Ct)],
                x4=[z(B, 9, 9, 9, 8 * 256), z(B, 9, 9, 9, 8 * 256)],
                h4=[z(B, 512, 512), z(B, 512, 512)],
                hd1=[z(NP, 1920), z(NP, 1920)], hd2=[z(NP, 768), z(NP, 768)], hd3=[z(NP, 384), z(NP, 384)],
                occ1=z(B, D ** 3, 8, dt=f32), occ2=z(B, D ** 3, 16, dt=f32),
                ws=z(3 * max(B * 4096 * 256, NP * 1920), dt=f32),
                bi=torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(P))
        b = self._wbufs[key]
        points = st["points"]

        def gemm3(A, W, bias, relu, M, N, K, out, *, mode=GEMM_LINEAR, lda=0, Do=0, Ci8=0,
                  out_mode=OUT_BF16, ldo=0, col_off=0):
            """out(hi, lo) = act(A.W^T + bias) from three bf16 GEMMs with fp32 slices."""
            for k, (a, ww) in enumerate(((A[0], W[0]), (A[1], W[0]), (A[0], W[1]))):
                sl = b["ws"][k * M * N:(k + 1) * M * N]
                self._gemm(L, a, ww, None, sl, M, N, K, mode=mode, lda=lda, Do=Do, Ci8=Ci8, relu=0,
                           out_mode=OUT_F32, ldo=N)
            chk(L.mf_px_combine(ptr(b["ws"]), 3, M, N, ptr(bias), relu, out_mode, Do, ptr(out[0]),
                                ptr(out[1]), ldo, col_off, s()), "px_combine")
            self.n_launches += 1

        with torch.cuda.device(dev):
            chk(L.mf_cnn_point_mlp_f32(
                ptr(st["values"]), ptr(points), ptr(w["conv1_rgb/W"]), ptr(w["conv1_rgb/b"]),
                ptr(w["conv1_pcd/W"]), ptr(w["conv1_pcd/b"]), ptr(w["conv2_rgb/W"]), ptr(w["conv2_rgb/b"]),
                ptr(w["conv2_pcd/W"]), ptr(w["conv2_pcd/b"]), B, P, D / 2.0 - 0.5, ptr(b["feat"][0]),
                FEAT_LD, ptr(b["feat2"]), ptr(b["feat1"]), s()), "point_mlp_f32")
            chk(L.mf_px_split(ptr(b["feat1"]), 72, NP, 72, ptr(b["feat"][0]), ptr(b["feat"][1]), FEAT_LD, 0, s()), "split")
            chk(L.mf_px_split(ptr(b["feat2"]), 144, NP, 144, ptr(b["feat"][0]), ptr(b["feat"][1]), FEAT_LD, 72, s()), "split")
            # _voxelize (model.py:143-164) through the public operator, occupancy stencils in fp32
            pts_np = points.permute(0, 2, 1).reshape(NP, 3).contiguous()
            with config.no_nan_check():
                vox, _ = AverageVoxelization3D.apply(b["feat2"], pts_np, b["bi"], B, (0.0, 0.0, 0.0), 1.0, (D, D, D))
            hocc = None
            if self._with_occupancy:
                g = st["gne"].to(torch.float32).contiguous()
                chk(L.mf_cnn_occ_convs(ptr(g), ptr(w["conv1_occ/W"]), ptr(w["conv1_occ/b"]),
                                       ptr(w["conv2_occ/W"]), ptr(w["conv2_occ/b"]), B, D, ptr(b["occ1"]),
                                       ptr(b["occ2"]), None, 0, 0, s()), "occ_convs")
                hocc = b["occ2"]
            chk(L.mf_px_pack_s2d(ptr(vox), ptr(hocc), B, 144, Cocc, D, ptr(b["x3"][0]), ptr(b["x3"][1]), s()), "px_pack")
            self.n_launches += 8
            gemm3(b["x3"], (px["conv3/hi"], px["conv3/lo"]), w["conv3/b"], 1, B * 4096, 256, 64 * Ct,
                  b["x4"], mode=GEMM_CONV_S2D, Do=16, Ci8=8 * Ct, out_mode=OUT_S2D_BF16)
            chk(L.mf_px_interp(ptr(b["x4"][0]), ptr(b["x4"][1]), 1, ptr(points), B, P, 256, 16, 2.0,
                               ptr(b["feat"][0]), ptr(b["feat"][1]), FEAT_LD, 216, s()), "px_interp3")
            gemm3(b["x4"], (px["conv4/hi"], px["conv4/lo"]), w["conv4/b"], 1, B * 512, 512, 64 * 256,
                  b["h4"], mode=GEMM_CONV_S2D, Do=8, Ci8=8 * 256, ldo=512)
            chk(L.mf_px_interp(ptr(b["h4"][0]), ptr(b["h4"][1]), 0, ptr(points), B, P, 512, 8, 4.0,
                               ptr(b["feat"][0]), ptr(b["feat"][1]), FEAT_LD, 472, s()), "px_interp4")
            gemm3(b["feat"], (px["head1/hi"], px["head1/lo"]), w["head1/b"], 1, NP, 1920, 984, b["hd1"],
                  lda=FEAT_LD, ldo=1920)
            heads = ("rot", "trans", "conf")
            for i, h in enumerate(heads):
                gemm3([t[:, i * 640:] for t in b["hd1"]], (px[f"conv2_{h}/hi"], px[f"conv2_{h}/lo"]),
                      w[f"conv2_{h}/b"], 1, NP, 256, 640, b["hd2"], lda=1920, ldo=768, col_off=i * 256)
            for i, h in enumerate(heads):
                gemm3([t[:, i * 256:] for t in b["hd2"]], (px[f"conv3_{h}/hi"], px[f"conv3_{h}/lo"]),
                      w[f"conv3_{h}/b"], 1, NP, 128, 256, b["hd3"], lda=768, ldo=384, col_off=i * 128)
            chk(L.mf_px_head4_pose(
                ptr(b["hd3"][0]), ptr(b["hd3"][1]), 384, ptr(px["conv4_rot/W32"]), ptr(w["conv4_rot/b"]),
                ptr(px["conv4_trans/W32"]), ptr(w["conv4_trans/b"]), ptr(px["conv4_conf/W32"]),
                ptr(w["conv4_conf/b"]), ptr(points), ptr(st["class_id"]), ptr(st["pitch"]), ptr(st["origin"]),
                B, P, nfg, ptr(out["rot"]), ptr(out["trans"]), ptr(out["conf"]), s()), "px_head4_pose")
            self.n_launches += 3

    # The forward is three fixed launch sequences (each CUDA-graph capturable):
    #   pre   : point MLP -> voxelise -> occupancy stencils -> s2d pack
    #   conv3 : the FLOP-dominant tcgen05 implicit GEMM (timed on its own by bench.py)
    #   post  : gather -> conv4 -> gather -> heads -> pose
    def _ctx(self, st):
        L = _lib.lib()
        dev = st["values"].device
        B, _, P = st["values"].shape
        # kernels read packed (bf16 / transposed) copies of the parameters: rebuild them whenever
        # any parameter changed (load_state_dict, optimizer step, in-place edit all bump the
        # tensor version) or moved
        ver = tuple(p._version for p in self.parameters())
        if (self._packed is None or self._packed_dev != self.conv3.weight.device
                or self._packed_ver != ver):
            self._pack()
            self._packed_ver = ver
        return L, dev, B, P, self._packed, self._work_buffers(B, P, dev)

    def _side(self, dev, which=0):
        key = (dev.type, dev.index, which)
        if key not in self._side_streams:
            self._side_streams[key] = torch.cuda.Stream(dev)
        return self._side_streams[key]

    def _occ_branch(self, L, st, w, buf, B, D, Ct):
        s = _lib.stream
        gne = st["gne"]
        fused = self.fused_occ
        if self.use_tensor_cores and D == 32 and gne.dtype in (torch.uint8, torch.bool):
            # byte grids go straight into the stencil (no cast pass)
            g = gne.contiguous()
            g = g.view(torch.uint8) if g.dtype == torch.bool else g
            if fused:
                _lib.check(L.mf_cnn_occ_fused_u8(
                    _lib.ptr(g), _lib.ptr(w["conv1_occ/W"]), _lib.ptr(w["conv1_occ/b"]),
                    _lib.ptr(w["conv2_occ/W"]), _lib.ptr(w["conv2_occ/b"]), B, D,
                    _lib.ptr(buf["x3"]), Ct, 144, s()), "occ_fused_u8")
                self.n_launches += 1
                return g
            _lib.check(L.mf_cnn_occ_convs_tc_u8(
                _lib.ptr(g), _lib.ptr(w["conv1_occ/W"]), _lib.ptr(w["conv1_occ/b"]),
                _lib.ptr(w["conv2_occ/W"]), _lib.ptr(w["conv2_occ/b"]), B, D,
                _lib.ptr(buf["occ1_bf16"]), _lib.ptr(buf["x3"]), Ct, 144, s()), "occ_convs_tc_u8")
            self.n_launches += 2
            return g
        g = gne.to(torch.float32).contiguous()
        if self.use_tensor_cores and D == 32 and fused:
            _lib.check(L.mf_cnn_occ_fused(
                _lib.ptr(g), _lib.ptr(w["conv1_occ/W"]), _lib.ptr(w["conv1_occ/b"]),
                _lib.ptr(w["conv2_occ/W"]), _lib.ptr(w["conv2_occ/b"]), B, D,
                _lib.ptr(buf["x3"]), Ct, 144, s()), "occ_fused")
            self.n_launches += 1
            return g
        if self.use_tensor_cores and D == 32:
            _lib.check(L.mf_cnn_occ_convs_tc(
                _lib.ptr(g), _lib.ptr(w["conv1_occ/W"]), _lib.ptr(w["conv1_occ/b"]),
                _lib.ptr(w["conv2_occ/W"]), _lib.ptr(w["conv2_occ/b"]), B, D,
                _lib.ptr(buf["occ1_bf16"]), _lib.ptr(buf["x3"]), Ct, 144, s()), "occ_convs_tc")
        else:
            _lib.check(L.mf_cnn_occ_convs(
                _lib.ptr(g), _lib.ptr(w["conv1_occ/W"]), _lib.ptr(w["conv1_occ/b"]),
                _lib.ptr(w["conv2_occ/W"]), _lib.ptr(w["conv2_occ/b"]), B, D,
                _lib.ptr(buf["occ1"]), None, _lib.ptr(buf["x3"]), Ct, 144, s()), "occ_convs")
        self.n_launches += 2
        return g

    def _stage_pre(self, st, pre_mlp=None):
        """pre_mlp: optional callable issued on the main stream after the side branches have been
        forked and before the point MLP (the e2e graph puts the H2D copy of `values` there, so the
        occupancy branch does not wait for it)."""
        L, dev, B, P, w, buf = self._ctx(st)
        D = self._voxel_dim
        s = _lib.stream
        with torch.cuda.device(dev):
            fused = self.fused_voxelize
            if fused and buf.get("x3_dense_dirty", False):
                buf["x3"].zero_()            # last call packed densely: sparse clear invalid
                buf["x3_dense_dirty"] = False
            vptr = _lib.ptr(st["values"])

            def point_mlp():
                _lib.check(L.mf_cnn_point_mlp(
                    vptr, _lib.ptr(st["points"]),
                    _lib.ptr(w["conv1_rgb/W"]), _lib.ptr(w["conv1_rgb/b"]),
                    _lib.ptr(w["conv1_pcd/W"]), _lib.ptr(w["conv1_pcd/b"]),
                    _lib.ptr(w["conv2_rgb/W"]), _lib.ptr(w["conv2_rgb/b"]),
                    _lib.ptr(w["conv2_pcd/W"]), _lib.ptr(w["conv2_pcd/b"]),
                    B, P, D / 2.0 - 0.5, _lib.ptr(buf["feat"]), FEAT_LD, _lib.ptr(buf["feat2"]),
                    s()), "point_mlp")

            forked = None
            forked2 = None
            keys_done = False
            if fused and self.concurrent_branches:
                # two independent branches: [point MLP + voxel bookkeeping (sparse clear of the
                # previous call's voxels, keys, sorted order) in one launch -> scatter] and the
                # occupancy stencils (they write channels [144,160) of x3, the point branch
                # channels [0,144): disjoint bytes).  The side stream forks from an event
                # recorded BEFORE the MLP launch, so it does not wait for it.
                main = torch.cuda.current_stream(dev)
                fork = torch.cuda.Event()
                fork.record(main)
                if pre_mlp is not None:
                    pre_mlp()
                    pre_mlp = None
                if P <= 4096:
                    _lib.check(L.mf_cnn_point_mlp_voxkeys(
                        vptr, _lib.ptr(st["points"]),
                        _lib.ptr(w["conv1_rgb/W"]), _lib.ptr(w["conv1_rgb/b"]),
                        _lib.ptr(w["conv1_pcd/W"]), _lib.ptr(w["conv1_pcd/b"]),
                        _lib.ptr(w["conv2_rgb/W"]), _lib.ptr(w["conv2_rgb/b"]),
                        _lib.ptr(w["conv2_pcd/W"]), _lib.ptr(w["conv2_pcd/b"]),
                        B, P, D / 2.0 - 0.5, _lib.ptr(buf["feat"]), FEAT_LD, _lib.ptr(buf["feat2"]),
                        144, D, 144 + (16 if self._with_occupancy else 0),
                        _lib.ptr(buf["prev_keys"]), _lib.ptr(buf["x3"]), s()), "point_mlp_voxkeys")
                    keys_done = True
                else:
                    side2 = self._side(dev, 1)
                    side2.wait_event(fork)
                    with torch.cuda.stream(side2):
                        _lib.check(L.mf_cnn_voxelize_s2d_phase(
                            None, _lib.ptr(st["points"]), B, P, 144, D,
                            144 + (16 if self._with_occupancy else 0), _lib.ptr(buf["prev_keys"]),
                            _lib.ptr(buf["x3"]), 1, s()), "voxelize_s2d(clear+keys)")
                    forked2 = side2
                    point_mlp()
                if self._with_occupancy:
                    side = self._side(dev)
                    side.wait_event(fork)
                    with torch.cuda.stream(side):
                        self._occ_branch(L, st, w, buf, B, D, 144 + 16)
                    forked = (main, side)
            else:
                if pre_mlp is not None:
                    pre_mlp()
                point_mlp()
            self.n_launches += 1
            Cocc = 16 if self._with_occupancy else 0
            Ct = 144 + Cocc
            g = (st["gne"].to(torch.float32).contiguous()
                 if (self._with_occupancy and not fused) else None)
            if fused:
                # _voxelize (model.py:143-164) fused with the bf16 s2d packing; the occupancy
                # stencil writes its 16 channels into the same buffer
                if forked2 is not None or keys_done:
                    if forked2 is not None:
                        torch.cuda.current_stream(dev).wait_stream(forked2)
                    _lib.check(L.mf_cnn_voxelize_s2d_phase(
                        _lib.ptr(buf["feat2"]), _lib.ptr(st["points"]), B, P, 144, D, Ct,
                        _lib.ptr(buf["prev_keys"]), _lib.ptr(buf["x3"]), 2, s()),
                        "voxelize_s2d(scatter)")
                else:
                    _lib.check(L.mf_cnn_voxelize_s2d(
                        _lib.ptr(buf["feat2"]), _lib.ptr(st["points"]), B, P, 144, D, Ct,
                        _lib.ptr(buf["prev_keys"]), _lib.ptr(buf["x3"]), s()), "voxelize_s2d")
                self.n_launches += 1 if keys_done else 3
                if forked:
                    forked[0].wait_stream(forked[1])
                elif self._with_occupancy:
                    self._occ_branch(L, st, w, buf, B, D, Ct)
            else:
                # unfused reference composition: the public operator + an explicit pack
                pts_np = st["points"].permute(0, 2, 1).reshape(B * P, 3).contiguous()
                from .... import config
                with config.no_nan_check():      # inputs are the model's own activations
                    vox, _ = AverageVoxelization3D.apply(
                        buf["feat2"], pts_np, buf["bi"], B, (0.0, 0.0, 0.0), 1.0, (D, D, D))
                self.n_launches += 2
                hocc = None
                if self._with_occupancy:
                    _lib.check(L.mf_cnn_occ_convs(
                        _lib.ptr(g), _lib.ptr(w["conv1_occ/W"]), _lib.ptr(w["conv1_occ/b"]),
                        _lib.ptr(w["conv2_occ/W"]), _lib.ptr(w["conv2_occ/b"]), B, D,
                        _lib.ptr(buf["occ1"]), _lib.ptr(buf["occ2"]), None, 0, 0, s()), "occ_convs")
                    self.n_launches += 2
                    hocc = buf["occ2"]
                _lib.check(L.mf_cnn_pack_s2d(_lib.ptr(vox), _lib.ptr(hocc), B, 144, Cocc, D,
                                             _lib.ptr(buf["x3"]), s()), "pack_s2d")
                self.n_launches += 1
                buf["prev_keys"].fill_(-1)
                buf["x3_dense_dirty"] = True

    def _stage_conv3(self, st):
        L, dev, B, P, w, buf = self._ctx(st)
        Ct = 144 + (16 if self._with_occupancy else 0)
        with torch.cuda.device(dev):
            # conv3: 32^3 x Ct -> 16^3 x 256, written straight into conv4's s2d input
            self._gemm(L, buf["x3"], w["conv3/W"], w["conv3/b"], buf["x4"], B * 4096, 256,
                       64 * Ct, mode=GEMM_CONV_S2D, Do=16, Ci8=8 * Ct, out_mode=OUT_S2D_BF16,
                       streamk=(buf["sk_sync"], buf["sk_ws"]))

    def _stage_post(self, st, out):
        L, dev, B, P, w, buf = self._ctx(st)
        NP = B * P
        nfg = self._n_fg_class
        points = st["points"]
        s = _lib.stream
        with torch.cuda.device(dev):
            def interp3():
                _lib.check(L.mf_cnn_interp_cl(_lib.ptr(buf["x4"]), 1, _lib.ptr(points), B, P, 256,
                                              16, 2.0, _lib.ptr(buf["feat"]), FEAT_LD, 216, s()),
                           "interp3")
            def conv4():
                # conv4: 16^3 x 256 -> 8^3 x 512
                self._gemm(L, buf["x4"], w["conv4/W"], w["conv4/b"], buf["h4"], B * 512, 512,
                           64 * 256, mode=GEMM_CONV_S2D, Do=8, Ci8=8 * 256, out_mode=OUT_BF16,
                           ldo=512, streamk=(buf["sk_sync"], buf["sk_ws"]))

            forked = None
            if self.concurrent_branches:
                # the conv3-level gather only reads x4: overlap it with conv4, whose split-K
                # launch leaves SMs idle.  conv4 is issued FIRST (its persistent CTAs need whole
                # SMs; behind the gather's 1000 CTAs it would start late) and the gather forks
                # from an event recorded before it.
                main, side = torch.cuda.current_stream(dev), self._side(dev)
                fork = torch.cuda.Event()
                fork.record(main)
                conv4()
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    interp3()
                forked = (main, side)
            else:
                interp3()
                conv4()
            _lib.check(L.mf_cnn_interp_cl(_lib.ptr(buf["h4"]), 0, _lib.ptr(points), B, P, 512, 8,
                                          4.0, _lib.ptr(buf["feat"]), FEAT_LD, 472, s()), "interp4")
            self.n_launches += 2
            if forked:
                forked[0].wait_stream(forked[1])
            # heads (model.py:239-254)
            heads = ("rot", "trans", "conf")
            if not (self.fused_heads and self.use_tensor_cores and self._heads_fused(L, buf, w, NP)):
                self._gemm(L, buf["feat"], w["head1/W"], w["head1/b"], buf["hd1"], NP, 1920, 984,
                           lda=FEAT_LD, ldo=1920)
                # layers 2-4 of the three heads: one grouped launch per layer (grid.z = head)
                self._gemm_grouped(L, [dict(
                    A=buf["hd1"][:, i * 640:], W=w[f"conv2_{h}/W"], bias=w[f"conv2_{h}/b"],
                    out=buf["hd2"], M=NP, N=256, K=640, lda=1920, ldo=768, col_off=i * 256)
                    for i, h in enumerate(heads)])
                self._gemm_grouped(L, [dict(
                    A=buf["hd2"][:, i * 256:], W=w[f"conv3_{h}/W"], bias=w[f"conv3_{h}/b"],
                    out=buf["hd3"], M=NP, N=128, K=256, lda=768, ldo=384, col_off=i * 128)
                    for i, h in enumerate(heads)])
            if self.fused_head4 and getattr(self, "_raw8", None) is not None:
                # training forward: also keep the 8 selected pre-activation outputs
                _lib.check(L.mf_cnn_head4_pose_train(
                    _lib.ptr(buf["hd3"]), 384,
                    _lib.ptr(w["conv4_rot/W"]), _lib.ptr(w["conv4_rot/b"]),
                    _lib.ptr(w["conv4_trans/W"]), _lib.ptr(w["conv4_trans/b"]),
                    _lib.ptr(w["conv4_conf/W"]), _lib.ptr(w["conv4_conf/b"]),
                    _lib.ptr(points), _lib.ptr(st["class_id"]), _lib.ptr(st["pitch"]),
                    _lib.ptr(st["origin"]), B, P, nfg,
                    _lib.ptr(out["rot"]), _lib.ptr(out["trans"]), _lib.ptr(out["conf"]),
                    _lib.ptr(self._raw8), s()), "head4_pose_train")
                self.n_launches += 1
                return
            if self.fused_head4:
                # layer 4 + class selection + pose epilogue in one kernel (only the object's
                # class rows are evaluated)
                _lib.check(L.mf_cnn_head4_pose(
                    _lib.ptr(buf["hd3"]), 384,
                    _lib.ptr(w["conv4_rot/W"]), _lib.ptr(w["conv4_rot/b"]),
                    _lib.ptr(w["conv4_trans/W"]), _lib.ptr(w["conv4_trans/b"]),
                    _lib.ptr(w["conv4_conf/W"]), _lib.ptr(w["conv4_conf/b"]),
                    _lib.ptr(points), _lib.ptr(st["class_id"]), _lib.ptr(st["pitch"]),
                    _lib.ptr(st["origin"]), B, P, nfg,
                    _lib.ptr(out["rot"]), _lib.ptr(out["trans"]), _lib.ptr(out["conf"]), s()),
                    "head4_pose")
                self.n_launches += 1
                return
            self._gemm_grouped(L, [dict(
                A=buf["hd3"][:, i * 128:], W=w[f"conv4_{h}/W"], bias=w[f"conv4_{h}/b"],
                out=buf["out_" + h], M=NP, N=buf["out_" + h].shape[1], K=128, lda=384, relu=0,
                out_mode=OUT_F32, ldo=buf["out_" + h].shape[1])
                for i, h in enumerate(heads)])
            _lib.check(L.mf_cnn_pose(
                _lib.ptr(buf["out_rot"]), _lib.ptr(buf["out_trans"]), _lib.ptr(buf["out_conf"]),
                _lib.ptr(points), _lib.ptr(st["class_id"]), _lib.ptr(st["pitch"]),
                _lib.ptr(st["origin"]), B, P, nfg,
                _lib.ptr(out["rot"]), _lib.ptr(out["trans"]), _lib.ptr(out["conf"]), s()), "pose")
            self.n_launches += 1

    # ------------------------------------------------------------------ resident runner
    def make_runner(self, B, P=1000, device=None, graph=True):
        return Runner(self, B, P, device or self.conv3.weight.device, graph)


class Runner:
    """Static device buffers + (optionally) three captured CUDA graphs for one batch shape:
    the public end-to-end call for serving: host batch in (pinned) -> poses out (pinned)."""

    def __init__(self, model, B, P, device, graph=True):
        self.model, self.B, self.P, self.device = model, B, P, device
        f32 = torch.float32
        # inputs and outputs live in ONE device blob / ONE pinned host blob each, so a step's
        # host->device and device->host traffic is a single copy either way
        # `values` (the big one, needed only by the point MLP) last: the e2e graph copies the
        # prefix first, starts the occupancy branch, and copies `values` while that runs
        in_spec = [("points", (B, 3, P), f32), ("class_id", (B,), torch.int32), ("pitch", (B,), f32),
                   ("origin", (B, 3), f32), ("gne", (B, 32, 32, 32), torch.uint8),
                   ("values", (B, 32, P), f32)]
        out_spec = [("rot", (B, P, 4), f32), ("trans", (B, P, 3), f32), ("conf", (B, P), f32)]
        self._in_spec, self._out_spec = in_spec, out_spec
        self.in_blob, self.st = self._blob(in_spec, device)
        self.out_blob, self.out = self._blob(out_spec, device)
        self.st["class_id"].fill_(1)
        self.st["pitch"].fill_(1.0)
        self.host_in_blob, self.host_in = self._blob(in_spec, "cpu")
        self.host_out_blob, self.host_out = self._blob(out_spec, "cpu")
        self.h2d_bytes = self.in_blob.numel()
        self.d2h_bytes = self.out_blob.numel()
        self.graphs = None
        self.ev = None
        self.launches_per_step = None
        if graph:
            self._capture()

    @staticmethod
    def _blob(spec, device):
        """One uint8 allocation with a 256-byte aligned typed view per entry."""
        offs, total = [], 0
        for _, shape, dt in spec:
            n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            offs.append(total)
            total += (n + 255) // 256 * 256
        blob = torch.zeros(total, dtype=torch.uint8, device=device)
        if str(device) == "cpu":
            blob = blob.pin_memory()
        views = {}
        for (name, shape, dt), o in zip(spec, offs):
            n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            views[name] = blob[o:o + n].view(dt).view(shape)
        return blob, views

    def new_host_blob(self):
        """A further pinned input blob with the same layout (for rotating host batches)."""
        return self._blob(self._in_spec, "cpu")

    def _eager(self):
        m = self.model
        n0 = m.n_launches
        m._stage_pre(self.st)
        if self.ev:
            self.ev[0].record()
        m._stage_conv3(self.st)
        if self.ev:
            self.ev[1].record()
        m._stage_post(self.st, self.out)
        self.launches_per_step = m.n_launches - n0

    def _capture(self):
        m = self.model
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._eager()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        gs = []
        for fn in (lambda: m._stage_pre(self.st), lambda: m._stage_conv3(self.st),
                   lambda: m._stage_post(self.st, self.out)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            gs.append(g)
        self.graphs = gs
        # the whole step as ONE graph (used whenever conv3 is not being timed on its own)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m._stage_pre(self.st)
            m._stage_conv3(self.st)
            m._stage_post(self.st, self.out)
        self.graph_step = g
        # the end-to-end step as ONE graph, from / to this runner's pinned blobs
        off = self.st["values"].data_ptr() - self.in_blob.data_ptr()
        self._values_off = off
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            # small inputs + occupancy grid first; `values` (1 MB, only the point MLP reads it) is
            # copied after the occupancy branch has been forked, i.e. while that branch runs
            self.in_blob[:off].copy_(self.host_in_blob[:off], non_blocking=True)
            m._stage_pre(self.st, pre_mlp=lambda: self.in_blob[off:].copy_(
                self.host_in_blob[off:], non_blocking=True))
            m._stage_conv3(self.st)
            # the poses are written by the last kernel straight into the pinned host output (UVA):
            # no D2H copy node and no kernel -> copy-engine hand-over at the end of the chain
            m._stage_post(self.st, self.host_out)
        self.graph_e2e = g

    def run_e2e(self):
        """Pinned host inputs (self.host_in) -> poses in pinned host memory (self.host_out): one
        graph launch containing both copies.  Fill self.host_in[...] first; the result is valid
        after a stream synchronisation."""
        if self.graphs is None:
            self.upload()
            self._eager()
            return self.download()
        self.graph_e2e.replay()
        return self.host_out

    def set_events(self, e0, e1):
        self.ev = (e0, e1)

    def run(self):
        if self.graphs is None:
            return self._eager()
        if not self.ev:
            return self.graph_step.replay()
        self.graphs[0].replay()
        if self.ev:
            self.ev[0].record()
        self.graphs[1].replay()
        if self.ev:
            self.ev[1].record()
        self.graphs[2].replay()

    def load_host(self, batch):
        """numpy/torch host batch -> pinned staging -> device (async on the current stream)."""
        for k, v in self.host_in.items():
            src = batch["grid_nontarget_empty" if k == "gne" else k]
            v.copy_(torch.as_tensor(src).to(v.dtype))
        self.upload()

    def upload(self, host_blob=None):
        """One async H2D copy of a pinned input blob (default: this runner's own)."""
        self.in_blob.copy_(self.host_in_blob if host_blob is None else host_blob,
                           non_blocking=True)

    def download(self):
        """One async D2H copy of rot/trans/conf into the pinned output blob."""
        self.host_out_blob.copy_(self.out_blob, non_blocking=True)
        return self.host_out
