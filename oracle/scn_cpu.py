"""``sparseconvnet``-shaped module surface on top of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Lets the reference's own ``3d_net/scn_unet.py`` (and the product's ``UNetSCN`` builder with
``backend=oracle.scn_cpu``) run on CPU with exactly the names, argument order, parameter
names and shapes SparseConvNet exposes.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import scn_oracle as O


class SparseConvNetTensor:
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def _ss(self):
        return int(self.spatial_size[0])


def _spatial(sz, dimension=3):
    if isinstance(sz, int):
        return torch.LongTensor([sz] * dimension)
    return torch.as_tensor(sz, dtype=torch.long)


class Sequential(nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, x):
        for m in self._modules.values():
            x = m(x)
        return x


class ConcatTable(nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, x):
        return [m(x) for m in self._modules.values()]


class Identity(nn.Module):
    def forward(self, x):
        return x


class JoinTable(nn.Module):
    def forward(self, xs):
        return SparseConvNetTensor(torch.cat([x.features for x in xs], 1), xs[0].metadata, xs[0].spatial_size)


class AddTable(nn.Module):
    def forward(self, xs):
        return SparseConvNetTensor(sum(x.features for x in xs), xs[0].metadata, xs[0].spatial_size)


class InputLayer(nn.Module):
    def __init__(self, dimension, spatial_size, mode=3):
        super().__init__()
        assert dimension == 3
        self.dimension = dimension
        self.spatial_size = _spatial(spatial_size, dimension)
        self.mode = mode

    def forward(self, x):
        coords, feats = x[0], x[1]
        meta = O.Metadata(coords.detach().cpu().long().numpy(), int(self.spatial_size[0]))
        return SparseConvNetTensor(O.input_layer(meta, feats.cpu(), self.mode), meta, self.spatial_size)


class OutputLayer(nn.Module):
    def __init__(self, dimension):
        super().__init__()

    def forward(self, x):
        return O.output_layer(x.metadata, x.features)


def _conv_weight(fv, n_in, n_out, groups=1):
    std = math.sqrt(2.0 / (n_in * fv))
    return nn.Parameter(torch.empty(fv, groups, n_in // groups, n_out // groups).normal_(0, std))


class SubmanifoldConvolution(nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, bias, groups=1):
        super().__init__()
        assert dimension == 3 and filter_size == 3 and groups == 1
        self.nIn, self.nOut = nIn, nOut
        self.weight = _conv_weight(27, nIn, nOut)
        if bias:
            self.bias = nn.Parameter(torch.zeros(nOut))

    def forward(self, x):
        y = O.submanifold_conv(x.metadata, x._ss(), x.features, self.weight)
        if hasattr(self, "bias"):
            y = y + self.bias
        return SparseConvNetTensor(y, x.metadata, x.spatial_size)


class Convolution(nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        assert dimension == 3 and filter_size == 2 and filter_stride == 2 and groups == 1
        self.weight = _conv_weight(8, nIn, nOut)
        if bias:
            self.bias = nn.Parameter(torch.zeros(nOut))

    def forward(self, x):
        y = O.conv_down(x.metadata, x._ss(), x.features, self.weight)
        if hasattr(self, "bias"):
            y = y + self.bias
        return SparseConvNetTensor(y, x.metadata, (x.spatial_size - 2) // 2 + 1)


class Deconvolution(nn.Module):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias, groups=1):
        super().__init__()
        assert dimension == 3 and filter_size == 2 and filter_stride == 2 and groups == 1
        self.weight = _conv_weight(8, nIn, nOut)
        if bias:
            self.bias = nn.Parameter(torch.zeros(nOut))

    def forward(self, x):
        out_ss = (x.spatial_size - 1) * 2 + 2
        y = O.deconv_up(x.metadata, int(out_ss[0]), x.features, self.weight)
        if hasattr(self, "bias"):
            y = y + self.bias
        return SparseConvNetTensor(y, x.metadata, out_ss)


class BatchNormLeakyReLU(nn.Module):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.leakiness = nPlanes, eps, momentum, leakiness
        self.weight = nn.Parameter(torch.ones(nPlanes))
        self.bias = nn.Parameter(torch.zeros(nPlanes))
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))

    def forward(self, x):
        y = O.batchnorm_relu(x.features, self.weight, self.bias, self.running_mean, self.running_var,
                             self.eps, self.momentum, self.training, self.leakiness)
        return SparseConvNetTensor(y, x.metadata, x.spatial_size)


class BatchNormReLU(BatchNormLeakyReLU):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, leakiness=0.0)


class NetworkInNetwork(nn.Module):
    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(nIn, nOut).normal_(0, math.sqrt(2.0 / nIn)))
        if bias:
            self.bias = nn.Parameter(torch.zeros(nOut))

    def forward(self, x):
        y = x.features @ self.weight
        if hasattr(self, "bias"):
            y = y + self.bias
        return SparseConvNetTensor(y, x.metadata, x.spatial_size)
