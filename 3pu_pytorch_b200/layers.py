"""Layer mirrors of the reference's network/layers.py: Conv1d (:161), Conv2d (:115), DenseEdgeConv (:6).

Constructor arguments, attribute names and therefore state_dict keys are the reference's
(`<name>.conv.weight`, `<name>.mlps.<i>.weight`, ...), so its checkpoints load unchanged.
The forward passes run on libpu3_b200 (fused kernels, see fused.py) when the input is an fp32
CUDA tensor and no normalisation layer is configured -- the only configuration the reference's
Level ever builds (upsampler.py:209-230).
"""
import torch
import torch.nn as nn

from . import fused
from .operations import group_knn


class DenseEdgeConv(nn.Module):
    """Dynamic-graph dense edge convolution (layers.py:6-64): kNN in feature space, edge feature
    [x_i, x_j - x_i], n 1x1 convolutions with dense concatenation ([new, old]), max over the k edges."""

    def __init__(self, in_channels, growth_rate, n, k, **kwargs):
        super(DenseEdgeConv, self).__init__()
        self.growth_rate = growth_rate
        self.n = n
        self.k = k
        self.mlps = torch.nn.ModuleList()
        self.mlps.append(torch.nn.Conv2d(2 * in_channels, growth_rate, 1, bias=True))
        for i in range(1, n):
            in_channels += growth_rate
            self.mlps.append(torch.nn.Conv2d(in_channels, growth_rate, 1, bias=True))

    def forward(self, x, idx=None):
        """x (B,C,N) -> y (B, C + n*growth_rate, N), idx (B,N,k) int64: ranks 1..k of the (k+1)-NN in
        feature space (rank 0 is dropped, not "self": layers.py:33-35)."""
        return fused.dense_edge_conv(x, [m.weight for m in self.mlps], [m.bias for m in self.mlps], self.k, idx)


class _ConvNd(nn.Module):
    """1x1 convolution with optional normalisation and activation (layers.py:115-204)."""
    _conv_cls = None
    _bn_cls = None
    _in_cls = None

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True,
                 activation=None, normalization=None, momentum=0.01):
        super(_ConvNd, self).__init__()
        self.activation = activation
        self.normalization = normalization
        bias = not normalization and bias
        self.conv = self._conv_cls(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=bias)
        if normalization is not None:
            if self.normalization == 'batch':
                self.norm = self._bn_cls(out_channels, affine=True, eps=0.001, momentum=momentum)
            elif self.normalization == 'instance':
                self.norm = self._in_cls(out_channels, affine=True, eps=0.001, momentum=momentum)
            else:
                raise ValueError("only \"batch/instance\" normalization permitted.")
        if activation is not None:
            if self.activation == 'relu':
                self.act = nn.ReLU()
            elif self.activation == 'elu':
                self.act = nn.ELU(alpha=1.0)
            elif self.activation == 'lrelu':
                self.act = nn.LeakyReLU(0.1)
            else:
                raise ValueError("only \"relu/elu/lrelu\" allowed")

    def _pointwise(self):
        ks = self.conv.kernel_size
        return all(k == 1 for k in ks) and all(s == 1 for s in self.conv.stride) and all(p == 0 for p in self.conv.padding)

    def forward(self, x, epoch=None):
        if self.normalization is None and self.activation in (None, 'relu') and self._pointwise() \
                and x.is_cuda and x.dtype == torch.float32:
            return fused.pointwise_conv(x, self.conv.weight, self.conv.bias, relu=self.activation == 'relu')
        # configurations the reference's Level never builds (norm layers, elu/lrelu, real kernels)
        x = self.conv(x)
        if self.normalization is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.act(x)
        return x


class Conv2d(_ConvNd):
    """2d convolution with custom normalization and activation (layers.py:115-158)."""
    _conv_cls, _bn_cls, _in_cls = nn.Conv2d, nn.BatchNorm2d, nn.InstanceNorm2d


class Conv1d(_ConvNd):
    """1d convolution with custom normalization and activation (layers.py:161-204)."""
    _conv_cls, _bn_cls, _in_cls = nn.Conv1d, nn.BatchNorm1d, nn.InstanceNorm1d
