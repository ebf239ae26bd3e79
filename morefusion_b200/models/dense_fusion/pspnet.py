"""PSPNet upsampler: 512-channel 1/8 features -> 32-channel per-pixel log-softmax features.

Architecture of morefusion/models/dense_fusion/pspnet.py:10-82: pyramid pooling at bin counts
1, 2, 3, 6 (average pooling with kernel = stride = H // bins, 1x1 conv without bias, bilinear
resize back with align_corners=True -- chainer's F.resize_images), bottleneck 1x1 conv + ReLU,
three x2 upsample blocks (resize, 3x3 conv, PReLU with one shared slope), dropout 0.3 / 0.15 /
0.15 in training, 1x1 conv to 32 channels, log-softmax over channels."""

import torch
import torch.nn.functional as F
from torch import nn


def _resize(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


class PSPModule(nn.Module):
    def __init__(self, in_channels, out_channels, sizes):
        super().__init__()
        for i in range(len(sizes)):
            setattr(self, f"conv{i + 1}", nn.Conv2d(in_channels, in_channels, 1, bias=False))
        self.bottleneck = nn.Conv2d(in_channels * (len(sizes) + 1), out_channels, 1)
        self.sizes = sizes

    def forward(self, x):
        H, W = x.shape[2:]
        hs = []
        for i, s in enumerate(self.sizes):
            k = (H // s, W // s)
            h = F.avg_pool2d(x, k, k)
            h = getattr(self, f"conv{i + 1}")(h)
            hs.append(_resize(h, (H, W)))
        hs.append(x)
        return torch.relu(self.bottleneck(torch.cat(hs, dim=1)))


class PSPUpsample(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, 3, 1, padding=1)
        self.prelu = nn.PReLU()                      # one shared slope, init 0.25 (L.PReLU())

    def forward(self, x):
        H, W = x.shape[2:]
        return self.prelu(self.conv(_resize(x, (H * 2, W * 2))))


class PSPNetExtractor(nn.Module):
    def __init__(self):
        super().__init__()
        self.psp = PSPModule(512, 1024, [1, 2, 3, 6])
        self.up1 = PSPUpsample(1024, 256)            # 1/8 -> 1/4
        self.up2 = PSPUpsample(256, 64)              # 1/4 -> 1/2
        self.up3 = PSPUpsample(64, 64)               # 1/2 -> 1
        self.conv1 = nn.Conv2d(64, 32, 1)

    def forward_up2(self, x):
        """Everything up to and including up2 (+ its dropout): [B,64,H/2,W/2]."""
        h = F.dropout(self.psp(x), 0.3, self.training)
        h = F.dropout(self.up1(h), 0.15, self.training)
        return F.dropout(self.up2(h), 0.15, self.training)

    def forward(self, x):
        h = self.up3(self.forward_up2(x))
        return F.log_softmax(self.conv1(h), dim=1)
