"""Slimmable (width-switchable) conv / BN -- drop-in for the reference's slimmable_ops.py.

Same class names, constructor signatures, attributes and state_dict keys as
search/slimmable_ops.py:5-70; both classes stay subclasses of nn.Conv2d / nn.BatchNorm2d because
tools/utils/init_func.py:5-15 and thop select modules by isinstance.  The arithmetic runs on the
sm_100a kernels of libfsb200 (no torch.nn.functional on the hot path).
"""
import torch
import torch.nn as nn

from . import engine
from . import functional as F_


def make_divisible(v, divisor=8, min_value=1):
    """Round a channel count to a multiple of `divisor`, never dropping by more than 10 %
    (reference: search/slimmable_ops.py:5-18)."""
    floor = divisor if min_value is None else min_value
    rounded = (int(v + divisor / 2) // divisor) * divisor
    out = rounded if rounded > floor else floor
    return out + divisor if out < 0.9 * v else out


class USConv2d(nn.Conv2d):
    """Conv2d whose active in/out channels are a ratio of the max width; the weight tensor keeps the max
    shape and the kernel reads the [:out, :in] corner in place (reference forward: slimmable_ops.py:36-48)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 depthwise=False, bias=True, width_mult_list=[1.]):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias)
        self.depthwise = depthwise
        self.in_channels_max = in_channels
        self.out_channels_max = out_channels
        self.width_mult_list = width_mult_list
        self.ratio = (1., 1.)

    def set_ratio(self, ratio):
        self.ratio = ratio

    def _resolve_channels(self):
        r_in, r_out = self.ratio
        assert r_in in self.width_mult_list, str(r_in) + " in? " + str(self.width_mult_list)
        assert r_out in self.width_mult_list, str(r_out) + " in? " + str(self.width_mult_list)
        # the reference mutates these attributes on every forward; callers (and thop) read them
        self.in_channels = make_divisible(self.in_channels_max * r_in)
        self.out_channels = make_divisible(self.out_channels_max * r_out)
        self.groups = self.in_channels if self.depthwise else 1
        return self.in_channels, self.out_channels

    def forward(self, input):
        return engine.conv_bn_act(input, self, None, relu=False)


class USBatchNorm2d(nn.BatchNorm2d):
    """One full nn.BatchNorm2d per candidate width under `self.bn` (keys `bn.{i}.*`); the module's own
    weight/bias (from super().__init__, track_running_stats=False) exist but are never used -- kept because
    they are part of the reference checkpoint format (slimmable_ops.py:51-62)."""

    def __init__(self, num_features, width_mult_list=[1.]):
        super().__init__(num_features, affine=True, track_running_stats=False)
        self.num_features_max = num_features
        self.width_mult_list = width_mult_list
        self.bn = nn.ModuleList(nn.BatchNorm2d(make_divisible(num_features * w), affine=True) for w in width_mult_list)
        self.ratio = 1.

    def set_ratio(self, ratio):
        self.ratio = ratio

    def _active_bn(self):
        assert self.ratio in self.width_mult_list
        return self.bn[self.width_mult_list.index(self.ratio)]

    def forward(self, input):
        return batchnorm_forward(input, self._active_bn())


def batchnorm_forward(x, bn, relu=False):
    """Stand-alone BatchNorm2d forward on our kernels (eval: folded affine; train: stats -> finalize -> apply)."""
    x = F_.to_nhwc_half(x)
    C = x.shape[1]
    if not bn.training and bn.running_mean is not None:
        scale, shift = engine.folded_bn(bn, C, None)
        return F_.affine_act(x, scale, shift, relu=relu)
    stats = engine.dp_allreduce_stats(F_.bn_stats(x))
    count = x.shape[0] * x.shape[2] * x.shape[3] * engine.dp_world_size()
    scale, shift, _, _ = F_.bn_finalize(stats, count, bn.weight, bn.bias, bn.eps,
                                        0.1 if bn.momentum is None else bn.momentum,
                                        bn.running_mean if bn.track_running_stats else None,
                                        bn.running_var if bn.track_running_stats else None)
    return F_.affine_act(x, scale, shift, relu=relu)
