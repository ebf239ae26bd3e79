"""Dilated ResNet-18 feature extractor, output stride 8, 512 channels.

Architecture of morefusion/models/resnet.py:7-52: chainercv2's ImageNet ResNet-18 with the
strides of stages 3 and 4 removed and their second units dilated by 2 and 4 (:24-35), batch
normalisation always in inference mode (:44), input normalised with the ImageNet mean / std
(:9-10, :43), and no gradient below res2 (``unchain`` :46-47).  Pretrained weights are a download
(chainercv2 model_provider, :18) and unavailable offline: parameters are randomly initialised and
loadable through ``load_state_dict``."""

import torch
from torch import nn


class _Unit(nn.Module):
    """chainercv2 ResUnit with a two-conv ResBlock body (conv3x3-BN-ReLU, conv3x3-BN)."""

    def __init__(self, cin, cout, stride, dilation):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, padding=dilation, dilation=dilation, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.identity_conv = None
        if cin != cout or stride != 1:
            self.identity_conv = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                               nn.BatchNorm2d(cout))

    def forward(self, x):
        identity = x if self.identity_conv is None else self.identity_conv(x)
        h = torch.relu(self.bn1(self.conv1(x)))
        h = self.bn2(self.conv2(h))
        return torch.relu(h + identity)


class _Stage(nn.Sequential):
    def __init__(self, cin, cout, stride, dilation):
        # unit1 keeps dilation 1 (only its stride is edited, resnet.py:24-25,30-31); unit2 dilated
        super().__init__(_Unit(cin, cout, stride, 1), _Unit(cout, cout, 1, dilation))


class ResNet18Extractor(nn.Module):

    mean_rgb = (0.485, 0.456, 0.406)
    std_rgb = (0.229, 0.224, 0.225)

    def __init__(self, unchain_at="res2"):
        assert unchain_at == "res2"
        super().__init__()
        self._unchain_at = unchain_at
        self.init_block = nn.Sequential(
            nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
            nn.MaxPool2d(3, 2, 1))
        self.res2 = _Stage(64, 64, 1, 1)
        self.res3 = _Stage(64, 128, 2, 1)
        self.res4 = _Stage(128, 256, 1, 2)          # stride removed, unit2 dilate 2
        self.res5 = _Stage(256, 512, 1, 4)          # stride removed, unit2 dilate 4
        self.register_buffer("mean", torch.tensor(self.mean_rgb)[None, :, None, None])
        self.register_buffer("std", torch.tensor(self.std_rgb)[None, :, None, None])

    def train(self, mode=True):
        # "disable update bn" (resnet.py:44): batch norm stays in inference mode
        super().train(mode)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        return self

    def forward(self, x):
        h = (x / 255.0 - self.mean) / self.std
        h = self.init_block(h)
        h = self.res2(h)
        if self._unchain_at == "res2":
            h = h.detach()
        h = self.res3(h)
        h = self.res4(h)
        return self.res5(h)
