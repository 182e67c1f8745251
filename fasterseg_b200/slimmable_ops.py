"""Slimmable (width-switchable) conv / BN -- drop-in for the reference's slimmable_ops.py.

Same class names, constructor signatures, attributes and state_dict keys as search/slimmable_ops.py:5-70; both classes stay
subclasses of nn.Conv2d / nn.BatchNorm2d because tools/utils/init_func.py:5-15 and thop select modules by isinstance.
The master tensors always have the maximum width; a width choice only selects which corner the kernels read
(`engine.conv_bn_act` asks `_resolve_channels()` / `_active_bn()`), so switching widths costs nothing and the packed fp16
weight copies are cached per (c_in, c_out) corner.  No torch.nn.functional on the hot path.
"""
import torch.nn as nn

from . import engine
from . import functional as F_


def make_divisible(v, divisor=8, min_value=1):
    """Channel count for a fractional width: nearest multiple of `divisor` (ties up), at least `min_value`
    (`divisor` when that is None), bumped by one `divisor` if rounding lost more than 10 % (slimmable_ops.py:5-18)."""
    lowest = divisor if min_value is None else min_value
    nearest = int(v + divisor / 2) // divisor * divisor
    channels = max(nearest, lowest)
    if channels < 0.9 * v:
        channels += divisor
    return channels


class _WidthSwitch:
    """`set_ratio` plumbing shared by the two slimmable modules: remembers the requested fraction(s) and validates them
    against the module's `width_mult_list` at use time (the reference asserts inside forward)."""

    def set_ratio(self, ratio):
        # plain attribute, written ~20 k times per supernet step: skip nn.Module.__setattr__'s Parameter / Module bookkeeping
        self.__dict__["ratio"] = ratio

    def _checked(self, fraction):
        assert fraction in self.width_mult_list, str(fraction) + " in? " + str(self.width_mult_list)
        return fraction


class USConv2d(_WidthSwitch, nn.Conv2d):
    """Conv2d over the active corner weight[:c_out, :c_in] of a max-width master weight (slimmable_ops.py:21-48)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 depthwise=False, bias=True, width_mult_list=[1.]):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                           groups=groups, bias=bias)
        self.in_channels_max, self.out_channels_max = in_channels, out_channels
        self.depthwise = depthwise
        self.width_mult_list = width_mult_list
        self.ratio = (1., 1.)

    def _resolve_channels(self):
        """(c_in, c_out) for the current ratio.  Like the reference's forward this also overwrites `in_channels`,
        `out_channels` and `groups`, which callers (and thop's counters) read afterwards."""
        ratio = self.ratio if type(self.ratio) is tuple else tuple(self.ratio)
        known = self.__dict__.setdefault("_corner_cache", {})
        active = known.get(ratio)
        if active is None:
            # a SymRatio (width chosen on the device, engine.SymRatio) runs the conv at its maximum width
            fractions = [1. if isinstance(r, engine.SymRatio) else self._checked(r) for r in ratio]
            active = tuple(make_divisible(full * r) for full, r in zip((self.in_channels_max, self.out_channels_max), fractions))
            known[ratio] = active
        state = self.__dict__     # in_channels / out_channels / groups are plain attributes of nn.Conv2d
        state["in_channels"], state["out_channels"] = active
        state["groups"] = active[0] if self.depthwise else 1
        return active

    def forward(self, input):
        return engine.conv_bn_act(input, self, None, relu=False)


class USBatchNorm2d(_WidthSwitch, nn.BatchNorm2d):
    """One complete nn.BatchNorm2d per candidate width under `self.bn` (state_dict keys `bn.{i}.*`).  The module's own
    `weight` / `bias` (created by the base class, no running statistics) are never used in forward but belong to the
    reference checkpoint format and to `parameters()` (slimmable_ops.py:51-62)."""

    def __init__(self, num_features, width_mult_list=[1.]):
        nn.BatchNorm2d.__init__(self, num_features, affine=True, track_running_stats=False)
        self.num_features_max = num_features
        self.width_mult_list = width_mult_list
        per_width = [nn.BatchNorm2d(make_divisible(num_features * fraction), affine=True) for fraction in width_mult_list]
        self.bn = nn.ModuleList(per_width)
        self.ratio = 1.

    def _active_bn(self):
        if isinstance(self.ratio, engine.SymRatio):
            ctx = engine.graph_ctx()
            assert ctx is not None, "a device-selected width needs an active graph context"
            return ctx.sel_bn(self, self.ratio.slot)
        return self.bn[self.width_mult_list.index(self._checked(self.ratio))]

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
