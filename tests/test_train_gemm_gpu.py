"""GPU: the training-side tcgen05 GEMMs (csrc/gemm_train.cu) against plain PyTorch fp32
references of the same operations (torch.matmul / torch.nn.grad.conv3d_*), bf16 operands."""

import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def L():
    from morefusion_b200 import _lib
    return _lib, _lib.lib()


def bf(x):
    return x.to(torch.bfloat16)


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("M,N,K,groups", [(8000, 1920, 984, 1), (8000, 256, 640, 3), (8000, 128, 256, 3),
                                          (500, 128, 64, 1)])
def test_gemm_tn(cuda_device, M, N, K, groups):
    """out[g][n, k] = sum_m dZ[m, g*N + n] * X[m, g*K + k]: both operands MN-major."""
    lib, l = L()
    torch.manual_seed(0)
    ldz = groups * N
    ldx = (groups * K + 63) // 64 * 64
    dZ = bf(torch.randn(M, ldz, device=cuda_device))
    X = bf(torch.randn(M, ldx, device=cuda_device))
    out = torch.full((groups, N, K), 7.0, device=cuda_device)
    rc = l.mf_train_gemm_tn(lib.ptr(dZ), ldz, lib.ptr(X), ldx, M, N, K, lib.ptr(out), K, groups,
                            N, K, N * K, 0, lib.stream())
    assert rc == 0
    for g in range(groups):
        want = dZ[:, g * N:(g + 1) * N].float().T @ X[:, g * K:(g + 1) * K].float()
        assert rel_err(out[g], want) < 2e-3, (g, rel_err(out[g], want))
    # accumulate = 1 adds on top
    rc = l.mf_train_gemm_tn(lib.ptr(dZ), ldz, lib.ptr(X), ldx, M, N, K, lib.ptr(out), K, groups,
                            N, K, N * K, 1, lib.stream())
    assert rc == 0
    want = 2 * (dZ[:, :N].float().T @ X[:, :K].float())
    assert rel_err(out[0], want) < 2e-3


def s2d_pack(x):
    """[B,C,D,D,D] fp32 -> bf16 [B,J,J,J,8*C] with J = D/2+1 (cnn.cu layout: xpad padded by 1)."""
    B, C, D = x.shape[:3]
    J = D // 2 + 1
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1, 1, 1))                  # [B,C,D+2,...]
    xp = xp.reshape(B, C, J, 2, J, 2, J, 2)                               # b c jd rd jh rh jw rw
    xp = xp.permute(0, 2, 4, 6, 3, 5, 7, 1)                               # b jd jh jw rd rh rw c
    return bf(xp.reshape(B, J, J, J, 8 * C)).contiguous()


def pad_cl(dy):
    """[B,C,Do,Do,Do] -> bf16 channels-last [B,Do+2,Do+2,Do+2,C] with a zero border."""
    return bf(torch.nn.functional.pad(dy, (1, 1, 1, 1, 1, 1)).permute(0, 2, 3, 4, 1)).contiguous()


@pytest.mark.parametrize("B,Ci,Co,D", [(2, 160, 256, 32), (2, 256, 512, 16)])
def test_conv_wgrad(cuda_device, B, Ci, Co, D):
    lib, l = L()
    from morefusion_b200.contrib.singleview_3d.models.model import pack_conv_k4s2_weight
    torch.manual_seed(1)
    Do = D // 2
    x = bf(torch.randn(B, Ci, D, D, D, device=cuda_device)).float()
    dy = bf(torch.randn(B, Co, Do, Do, Do, device=cuda_device)).float()
    want = torch.nn.grad.conv3d_weight(x, (Co, Ci, 4, 4, 4), dy, stride=2, padding=1)
    want = pack_conv_k4s2_weight(want).float()                            # [Co, 64*Ci], tap-major
    out = torch.full((Co, 64 * Ci), 3.0, device=cuda_device)
    dyp, xs = pad_cl(dy), s2d_pack(x)                 # keep the operands alive across the launch
    rc = l.mf_train_conv_wgrad(lib.ptr(dyp), lib.ptr(xs), B, Do, Co, 8 * Ci,
                               lib.ptr(out), 0, lib.stream())
    assert rc == 0
    torch.cuda.synchronize()
    assert rel_err(out, want) < 1e-2, rel_err(out, want)                  # `want` is bf16-rounded


def pack_dgrad_weight(W):
    """OIDHW [Co,Ci,4,4,4] -> bf16 [8][Ci][8*Co]: Wd[r][ci][a*Co+co] = W[co][ci][2a+r]."""
    Co, Ci = W.shape[:2]
    W = W.reshape(Co, Ci, 2, 2, 2, 2, 2, 2)                # co ci ad rd ah rh aw rw
    W = W.permute(3, 5, 7, 1, 2, 4, 6, 0)                  # rd rh rw ci ad ah aw co
    return bf(W.reshape(8, Ci, 8 * Co)).contiguous()


@pytest.mark.parametrize("B,Ci,Co,D,epi", [(2, 256, 512, 16, 1), (2, 160, 256, 32, 2)])
def test_conv_dgrad(cuda_device, B, Ci, Co, D, epi):
    lib, l = L()
    torch.manual_seed(2)
    Do = D // 2
    W = bf(torch.randn(Co, Ci, 4, 4, 4, device=cuda_device) / 30).float()
    dy = bf(torch.randn(B, Co, Do, Do, Do, device=cuda_device)).float()
    want = torch.nn.grad.conv3d_input((B, Ci, D, D, D), W, dy, stride=2, padding=1)
    want_cl = want.permute(0, 2, 3, 4, 1).contiguous()                    # [B,D,D,D,Ci]
    Wd = pack_dgrad_weight(W)
    dyp = pad_cl(dy)                                  # keep alive across the launches
    if epi == 1:
        out = torch.ones(B, D, D, D, Ci, device=cuda_device)               # accumulated into
        rc = l.mf_train_conv_dgrad(lib.ptr(dyp), lib.ptr(Wd), B, Do, Co, Ci, 1, lib.ptr(out),
                                   Ci, 0, lib.stream())
        assert rc == 0
        assert rel_err(out - 1.0, want_cl) < 1e-2
    else:
        M = B * Do ** 3
        out = torch.zeros(8, M, Ci, dtype=torch.bfloat16, device=cuda_device)
        rc = l.mf_train_conv_dgrad(lib.ptr(dyp), lib.ptr(Wd), B, Do, Co, Ci, 2, lib.ptr(out),
                                   Ci, M * Ci, lib.stream())
        assert rc == 0
        got = torch.zeros_like(want_cl)
        o = out.float().reshape(2, 2, 2, B, Do, Do, Do, Ci)
        for rd in range(2):
            for rh in range(2):
                for rw in range(2):
                    got[:, 1 - rd::2, 1 - rh::2, 1 - rw::2] = o[rd, rh, rw]    # x = 2 o + 1 - r
        assert rel_err(got, want_cl) < 1.5e-2
