"""Host-side check of the space-to-depth (s2d) GEMM formulation of the k4/s2/p1 Conv3D that
`csrc/conv3d_tc.cu` implements (forward) and that the training step will reuse (weight and
input gradients): every identity below is evaluated with plain torch-CPU matmuls on the
repo's own packing helpers and compared with torch's conv3d / autograd.

Layouts:
  xpad = zero-pad(X, 1)                      [B, C, 2J, 2J, 2J],  J = Do + 1, Do = D / 2
  Xs[b, jd, jh, jw, r*C + ci] = xpad[b, ci, 2jd+rd, 2jh+rh, 2jw+rw],   r = (rd*2 + rh)*2 + rw
  Wg[co, a*8C + r*C + ci]     = W[co, ci, 2ad+rd, 2ah+rh, 2aw+rw],     a = (ad*2 + ah)*2 + aw
forward   Y[(b,o), co]      = sum_a  Xs[b, o + a, :] . Wg[co, a*8C:(a+1)*8C]
wgrad     dWg[co, a*8C + k] = sum_(b,o) dY[(b,o), co] * Xs[b, o + a, k]
dgrad     dXs[b, j, k]      = sum_a  dYp[b, j - a + 1, :] . Wg[:, a*8C + k]   (dYp = dY zero-padded by 1)
"""

import pytest
import torch

from morefusion_b200.contrib.singleview_3d.models.model import pack_conv_k4s2_weight


def s2d(x):
    """[B,C,D,D,D] -> [B,J,J,J,8C] (fp32; the CUDA path stores bf16)."""
    B, C, D = x.shape[:3]
    J = D // 2 + 1
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1, 1, 1))
    return xp.reshape(B, C, J, 2, J, 2, J, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(B, J, J, J, 8 * C)


def un_s2d(xs, C):
    """inverse of s2d, dropping the padding ring: [B,J,J,J,8C] -> [B,C,D,D,D]."""
    B, J = xs.shape[:2]
    xp = xs.reshape(B, J, J, J, 2, 2, 2, C).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(B, C, 2 * J, 2 * J, 2 * J)
    return xp[:, :, 1:-1, 1:-1, 1:-1]


def offsets():
    return [(ad, ah, aw) for ad in (0, 1) for ah in (0, 1) for aw in (0, 1)]


@pytest.mark.parametrize("B,C,Co,D", [(2, 6, 5, 8), (1, 4, 7, 12)])
def test_s2d_forward_wgrad_dgrad_identities(B, C, Co, D):
    torch.manual_seed(B * 100 + C)
    Do, J = D // 2, D // 2 + 1
    x = torch.randn(B, C, D, D, D, dtype=torch.float64, requires_grad=True)
    W = torch.randn(Co, C, 4, 4, 4, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv3d(x, W, stride=2, padding=1)                 # [B,Co,Do,Do,Do]
    gy = torch.randn_like(y)
    gx_ref, gW_ref = torch.autograd.grad(y, (x, W), gy)

    Xs = s2d(x.detach())
    # the product's packer emits bf16 and its index order is what is under test: push the flat
    # source indices through it as two base-256 digits (exact in bf16) to recover the permutation
    idx = torch.arange(Co * C * 64, dtype=torch.float32).reshape(Co, C, 4, 4, 4)
    lo = pack_conv_k4s2_weight((idx % 256)).float()
    hi = pack_conv_k4s2_weight(torch.div(idx, 256, rounding_mode="floor")).float()
    perm = (hi * 256 + lo).long()                                             # [Co, 64C] flat source index
    assert perm.max() < 65536, "test sizes must keep the index digits bf16-exact"
    Wg = W.detach().reshape(-1)[perm]                                         # [Co, 64*C]
    K8 = 8 * C

    # ---- forward
    Y = torch.zeros(B, Do, Do, Do, Co, dtype=torch.float64)
    for a, (ad, ah, aw) in enumerate(offsets()):
        A = Xs[:, ad:ad + Do, ah:ah + Do, aw:aw + Do, :]                      # [B,Do,Do,Do,8C]
        Y += A @ Wg[:, a * K8:(a + 1) * K8].T
    torch.testing.assert_close(Y.permute(0, 4, 1, 2, 3), y.detach(), rtol=1e-12, atol=1e-12)

    # ---- wgrad: reduction over the M = B*Do^3 output rows
    gY = gy.permute(0, 2, 3, 4, 1).reshape(-1, Co)                            # [M, Co]
    gWg = torch.zeros(Co, 64 * C, dtype=torch.float64)
    for a, (ad, ah, aw) in enumerate(offsets()):
        A = Xs[:, ad:ad + Do, ah:ah + Do, aw:aw + Do, :].reshape(-1, K8)      # [M, 8C]
        gWg[:, a * K8:(a + 1) * K8] = gY.T @ A
    gW = torch.zeros(Co * C * 64, dtype=torch.float64)
    gW[perm.reshape(-1)] = gWg.reshape(-1)                                    # un-pack
    torch.testing.assert_close(gW.reshape(Co, C, 4, 4, 4), gW_ref, rtol=1e-11, atol=1e-11)

    # ---- dgrad: k2/s1 conv over the zero-padded dY with flipped cell offsets, then un-s2d
    gYp = torch.nn.functional.pad(gy.permute(0, 2, 3, 4, 1), (0, 0, 1, 1, 1, 1, 1, 1))   # [B,Do+2,..,Co]
    gXs = torch.zeros(B, J, J, J, K8, dtype=torch.float64)
    for a, (ad, ah, aw) in enumerate(offsets()):
        G = gYp[:, 1 - ad:1 - ad + J, 1 - ah:1 - ah + J, 1 - aw:1 - aw + J, :]           # dYp[j - a + 1]
        gXs += G @ Wg[:, a * K8:(a + 1) * K8]
    torch.testing.assert_close(un_s2d(gXs, C), gx_ref, rtol=1e-11, atol=1e-11)
    # the padding ring of dXs receives gradient too (it multiplies zeros in the forward pass):
    # the CUDA path must drop it, which un_s2d does by construction
