"""Stand-in for the reference's 2D branch (``Net2DSeg``, ``2d_net/model.py:35-142`` + ``backbones.py``) for the
full-step benchmark (BASELINE configs[2]): the same layer types and sizes on stock PyTorch / cuDNN, nothing custom --
SURVEY.md 8 keeps the 2D network out of the hot path ("dense cuDNN work used as-is"); only its lift is ours.

Shape of the network: two ResNet-34 encoders (RGB 3-channel, sparse depth 1-channel) whose 7x7 stem has stride 1, five
feature scales (64, 64, 128, 256, 512 channels at 1, 1/2, 1/4, 1/8, 1/16), dropout 0.4 after the two deepest stages; a
decoder that upsamples with 2x2 stride-2 transposed convolutions (+BN+ReLU) and fuses [depth skip | upsampled | RGB skip]
with 3x3 convolutions (+BN+ReLU); a 64-channel full-resolution feature map; two heads of 5x5 average pooling + 1x1
convolution to ``num_classes`` logits, both lifted to the LiDAR points with ``mm2d3d_b200.lift.lift2d``.  Inputs are
padded to a multiple of 16 and the feature map is cropped back (225x400 -> 240x400 -> 225x400).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models.resnet import resnet34

from mm2d3d_b200.lift import lift2d

SCALES = (64, 64, 128, 256, 512)


class _Encoder(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        r = resnet34(weights=None)
        self.stem = nn.Sequential(nn.Conv2d(in_channels, 64, 7, 1, 3, bias=False), r.bn1, nn.ReLU(inplace=True))
        self.stages = nn.ModuleList([nn.Sequential(r.maxpool, r.layer1), r.layer2, r.layer3, r.layer4])
        self.drop = nn.Dropout(0.4)

    def forward(self, x):
        feats = [self.stem(x)]
        for i, st in enumerate(self.stages):
            x = st(feats[-1])
            feats.append(self.drop(x) if i >= 2 else x)
        return feats


def _up(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 2, 2), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


def _fuse(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class RGBDUNet2D(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.rgb, self.depth = _Encoder(3), _Encoder(1)
        # deepest first: (upsample in -> out, fuse 3*out -> out)
        self.ups = nn.ModuleList([_up(2 * SCALES[4], SCALES[3]), _up(SCALES[3], SCALES[2]), _up(SCALES[2], SCALES[1]),
                                  _up(SCALES[1], SCALES[0])])
        self.fuses = nn.ModuleList([_fuse(3 * SCALES[3], SCALES[3]), _fuse(3 * SCALES[2], SCALES[2]),
                                    _fuse(3 * SCALES[1], SCALES[1]), nn.Conv2d(3 * SCALES[0], 64, 3, padding=1)])
        self.pool = nn.AvgPool2d(5, 1, 2)
        self.cls, self.cls_aux = nn.Conv2d(64, num_classes, 1), nn.Conv2d(64, num_classes, 1)

    def dense(self, img, depth):
        """The dense part (static shapes: what a CUDA graph can hold): logits of both heads and the feature map."""
        h, w = img.shape[-2:]
        ph, pw = (-h) % 16, (-w) % 16
        if ph or pw:
            img, depth = F.pad(img, [0, pw, 0, ph]), F.pad(depth, [0, pw, 0, ph])
        a, b = self.rgb(img), self.depth(depth)
        x = torch.cat([b[4], a[4]], 1)
        for i, (up, fuse) in enumerate(zip(self.ups, self.fuses)):
            s = 3 - i
            x = fuse(torch.cat([b[s], up(x), a[s]], 1))
        fmap = x[:, :, :h, :w]
        pooled = self.pool(fmap)
        return self.cls(pooled), self.cls_aux(pooled), fmap

    def forward(self, img, depth, lift_indices):
        logit_2d, logit_aux_2d, fmap = self.dense(img, depth)
        # the lift: one gather kernel over all samples instead of the reference's per-sample indexing loop
        return lift2d(logit_2d.float(), lift_indices), lift2d(logit_aux_2d.float(), lift_indices), fmap


class Dense2D(nn.Module):
    """``RGBDUNet2D.dense`` as a module of its own (for ``torch.cuda.make_graphed_callables``)."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, img, depth):
        a, b, _ = self.net.dense(img, depth)
        return a, b
