"""Segmentation head operators -- drop-in for the live part of the reference's seg_oprs.py
(ConvBnRelu :17-39, FeatureFusion :181-225, Head :228-274; the rest of that file is dead code).
Parameter names (`conv.weight`, `bn.*`, `conv_3x3.*`, `conv_1x1.{weight,bias}`, the unused
`channel_attention.{1,2}.conv.weight`) are the checkpoint format and are kept."""
import os.path as osp

import numpy as np
import torch.nn as nn

from . import engine
from .operations import _table_latency, compute_latency


class ConvBnRelu(nn.Module):
    def __init__(self, in_planes, out_planes, ksize, stride, pad, dilation=1, groups=1, has_bn=True,
                 norm_layer=nn.BatchNorm2d, bn_eps=1e-5, has_relu=True, inplace=True, has_bias=False):
        super(ConvBnRelu, self).__init__()
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=ksize, stride=stride, padding=pad, dilation=dilation,
                              groups=groups, bias=has_bias)
        self.has_bn = has_bn
        if has_bn:
            self.bn = norm_layer(out_planes, eps=bn_eps)
        self.has_relu = has_relu
        if has_relu:
            self.relu = nn.ReLU(inplace=inplace)

    def forward(self, x, out=None):
        # conv (+bias) -> BN -> ReLU fused into one launch
        return engine.conv_bn_act(x, self.conv, self.bn if self.has_bn else None, relu=self.has_relu, out=out)


class FeatureFusion(nn.Module):
    """forward == one 1x1 ConvBnRelu on the already-concatenated branch features (seg_oprs.py:219-222); the
    channel-attention branch is commented out upstream but its parameters are part of the state_dict."""

    def __init__(self, in_planes, out_planes, reduction=1, Fch=16, scale=4, branch=2, norm_layer=nn.BatchNorm2d):
        super(FeatureFusion, self).__init__()
        self.conv_1x1 = ConvBnRelu(in_planes, out_planes, 1, 1, 0, has_bn=True, norm_layer=norm_layer, has_relu=True,
                                   has_bias=False)
        self.channel_attention = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            ConvBnRelu(out_planes, out_planes // reduction, 1, 1, 0, has_bn=False, norm_layer=norm_layer, has_relu=True,
                       has_bias=False),
            ConvBnRelu(out_planes // reduction, out_planes, 1, 1, 0, has_bn=False, norm_layer=norm_layer, has_relu=False,
                       has_bias=False),
            nn.Sigmoid())
        self._Fch, self._scale, self._branch = Fch, scale, branch

    @staticmethod
    def _latency(h, w, C_in, C_out):
        return compute_latency(FeatureFusion(C_in, C_out), (1, C_in, h, w))

    def forward_latency(self, size):
        name = "ff_H%d_W%d_C%d" % (size[1], size[2], size[0])
        c = self._scale * self._Fch * self._branch
        return _table_latency(name, lambda: FeatureFusion._latency(size[1], size[2], c, c)), size

    def forward(self, fm, out=None):
        return self.conv_1x1(fm, out=out)


class Head(nn.Module):
    """3x3 ConvBnRelu (mid = in if in <= 256 else in // 2) -> 1x1 conv WITH bias to the class logits
    (seg_oprs.py:228-274).  The bias is folded into the 1x1 kernel's epilogue shift."""

    def __init__(self, in_planes, out_planes=19, Fch=16, scale=4, branch=2, is_aux=False, norm_layer=nn.BatchNorm2d):
        super(Head, self).__init__()
        mid_planes = in_planes if in_planes <= 256 else in_planes // 2
        self.conv_3x3 = ConvBnRelu(in_planes, mid_planes, 3, 1, 1, has_bn=True, norm_layer=norm_layer, has_relu=True,
                                   has_bias=False)
        self.conv_1x1 = nn.Conv2d(mid_planes, out_planes, kernel_size=1, stride=1, padding=0)
        self._in_planes, self._out_planes = in_planes, out_planes
        self._Fch, self._scale, self._branch = Fch, scale, branch

    @staticmethod
    def _latency(h, w, C_in, C_out=19):
        return compute_latency(Head(C_in, C_out), (1, C_in, h, w))

    def forward_latency(self, size):
        assert size[0] == self._in_planes, "size[0] %d, self._in_planes %d" % (size[0], self._in_planes)
        name = "head_H%d_W%d_Cin%d_Cout%d" % (size[1], size[2], size[0], self._out_planes)
        latency = _table_latency(name, lambda: Head._latency(size[1], size[2], self._scale * self._Fch * self._branch,
                                                             self._out_planes))
        return latency, (self._out_planes, size[1], size[2])

    def forward(self, x, out=None):
        fm = self.conv_3x3(x)
        return engine.conv_bn_act(fm, self.conv_1x1, None, relu=False, out=out)
