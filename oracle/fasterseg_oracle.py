"""TEST INFRASTRUCTURE ONLY -- CPU fp32 oracle for the FasterSeg conv hot path.

A from-scratch restatement (torch CPU fp32, NCHW, *functional*: no nn.Module, parameters
come from a state_dict that uses the reference's key names) of the reference algorithm for
the path SURVEY.md section 8 scopes.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this file; the product
package `fasterseg_b200/` never does.

Pinned against the reference's own modules by `oracle/make_golden.py` (run where
/root/reference exists) -> `tests/golden/*.npz|json`, checked by
`tests/test_oracle_golden.py`.  The reference ships no tests of its own for this path
(SURVEY.md section 4), so the pins are outputs of the reference itself run in the build
container plus the three reference-shipped known answers (genotype decode of arch_1.pt,
latency12/latency02, make_divisible).

All `file:line` citations are relative to the reference tree root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

PRIMITIVES = ["skip", "conv", "conv_downup", "conv_2x", "conv_2x_downup"]  # search/genotypes.py:5-11
BN_EPS = 1e-5       # search/config_search.py:52 (set on every BN by tools/utils/init_func.py:10-13)
BN_MOMENTUM = 0.1   # search/config_search.py:53
WIDTH_MULT_LIST = [4. / 12, 6. / 12, 8. / 12, 10. / 12, 1.]  # search/config_search.py:83


# --------------------------------------------------------------------------------------
# scalar helpers
# --------------------------------------------------------------------------------------
def make_divisible(v, divisor=8, min_value=1):
    """search/slimmable_ops.py:5-18."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def num_filters(scale, Fch=12, width=1.0):
    """train/model_seg.py:205-206, search/model_search.py:200-201."""
    return int(np.round(scale * Fch * width))


def conv_padding(kernel_size, stride, dilation=1):
    """ConvNorm's default padding, search/operations.py:55-58."""
    return int(np.ceil((dilation * (kernel_size - 1) + 1 - stride) / 2.))


# --------------------------------------------------------------------------------------
# optional fp16-storage emulation: with EMULATE_FP16["on"] every activation an fp16-storage implementation would write to
# memory (unit outputs, resize outputs, the image, the weights) is rounded to fp16 in the forward pass, and the matching
# activation gradients are rounded (at loss scale EMULATE_FP16["gscale"]) in the backward pass.  Arithmetic stays fp32.
# Used by the tests to separate "implementation error" from the conditioning of train-mode BatchNorm chains: the CUDA path
# must be as close to the fp32 oracle as this emulation is.
# --------------------------------------------------------------------------------------
EMULATE_FP16 = {"on": False, "gscale": 1024.0}


def _q(t):
    """Round an activation to fp16 storage (straight-through) + round its gradient, when emulation is on."""
    if not EMULATE_FP16["on"]:
        return t
    out = t + (t.detach().half().float() - t.detach())
    if out.requires_grad:
        s = EMULATE_FP16["gscale"]
        out.register_hook(lambda g: (g * s).half().float() / s)
    return out


def _qw(w):
    if not EMULATE_FP16["on"]:
        return w
    return w + (w.detach().half().float() - w.detach())


# --------------------------------------------------------------------------------------
# tensor primitives (what L0 = torch.nn.functional does for the reference)
# --------------------------------------------------------------------------------------
def conv2d(x, w, bias=None, stride=1, padding=0, dilation=1):
    """F.conv2d call site search/slimmable_ops.py:47 and every nn.Conv2d in operations.py."""
    return F.conv2d(_q(x) if x.shape[1] == 3 else x, _qw(w), bias, stride, padding, dilation, 1)


def batchnorm(x, weight, bias, running_mean, running_var, training, momentum=BN_MOMENTUM, eps=BN_EPS):
    """nn.BatchNorm2d semantics used everywhere in operations.py / seg_oprs.py.

    eval : y = (x - running_mean) / sqrt(running_var + eps) * weight + bias
    train: batch mean / *biased* var for normalisation; running stats updated in place with
           momentum and the *unbiased* var (SURVEY.md section 7 "BN semantics").
    Returns y; updates running_mean / running_var in place when training (like the module).
    """
    if not training:
        scale = weight / torch.sqrt(running_var + eps)
        shift = bias - running_mean * scale
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    n = x.numel() // x.size(1)
    mean = x.mean(dim=(0, 2, 3))
    var_b = ((x - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
    if running_mean is not None:
        with torch.no_grad():
            var_u = var_b * (n / max(n - 1, 1))
            running_mean.mul_(1 - momentum).add_(momentum * mean.detach())
            running_var.mul_(1 - momentum).add_(momentum * var_u.detach())
    y = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var_b.view(1, -1, 1, 1) + eps)
    return y * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)


RESIZE_IMPL = {"aten": False}  # bench.py's CPU-baseline legs switch to the reference's own ATen call for a fair timing


def bilinear_ac(x, size):
    """F.interpolate(mode='bilinear', align_corners=True), search/operations.py:271,275,437,444.

    src = dst * (in - 1) / (out - 1) (0 when out == 1); two-tap lerp per axis.
    Written out (not via F.interpolate) so that the oracle states the formula itself; tests/test_oracle_golden.py checks the
    formula against F.interpolate outputs of the reference.  RESIZE_IMPL["aten"] routes to F.interpolate (what the
    reference executes) -- used only when TIMING the CPU baseline, where the gather-based formula would be unfairly slow.
    """
    N, C, H, W = x.shape
    Ho, Wo = int(size[0]), int(size[1])
    if RESIZE_IMPL["aten"]:
        return F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=True)

    def taps(n_in, n_out):
        if n_out > 1:
            scale = (n_in - 1) / (n_out - 1)
        else:
            scale = 0.0
        src = torch.arange(n_out, dtype=torch.float32) * torch.tensor(scale, dtype=torch.float32)
        i0 = src.floor().to(torch.int64).clamp_(0, n_in - 1)
        i1 = (i0 + 1).clamp_(max=n_in - 1)
        l1 = src - i0.to(torch.float32)
        return i0, i1, l1

    h0, h1, lh = taps(H, Ho)
    w0, w1, lw = taps(W, Wo)
    top = x[:, :, h0, :]
    bot = x[:, :, h1, :]
    lh = lh.view(1, 1, Ho, 1)
    rows = top + (bot - top) * lh  # same association order does not matter at fp32 tolerance
    left = rows[:, :, :, w0]
    right = rows[:, :, :, w1]
    lw = lw.view(1, 1, 1, Wo)
    return left + (right - left) * lw


# --------------------------------------------------------------------------------------
# parameter access by reference key names
# --------------------------------------------------------------------------------------
class Params:
    """Thin view over a state_dict with a key prefix."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        self.sd = sd
        self.prefix = prefix

    def sub(self, name: str) -> "Params":
        return Params(self.sd, self.prefix + name + ".")

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.sd[self.prefix + name]

    def has(self, name: str) -> bool:
        return (self.prefix + name) in self.sd


def _bn(x, p: Params, training: bool):
    return batchnorm(x, p["weight"], p["bias"], p["running_mean"], p["running_var"], training)


def _us_bn(x, p: Params, width_idx: Optional[int], training: bool):
    """USBatchNorm2d dispatch: one nn.BatchNorm2d per width, search/slimmable_ops.py:51-70."""
    if width_idx is None:
        return _bn(x, p, training)
    return _bn(x, p.sub("bn.%d" % width_idx), training)


def _us_w(w, ci, co):
    """USConv2d weight slice, search/slimmable_ops.py:42."""
    return w[:co, :ci]


class Ratio:
    """Resolved (ratio_in, ratio_out) for one slimmable op, search/operations.py set_ratio."""

    def __init__(self, r_in: float, r_out: float, width_mult_list: Sequence[float]):
        self.r_in, self.r_out = r_in, r_out
        self.wml = list(width_mult_list)

    def ci(self, c_max):
        return make_divisible(c_max * self.r_in)

    def co(self, c_max):
        return make_divisible(c_max * self.r_out)

    def idx_out(self):
        return self.wml.index(self.r_out)


# --------------------------------------------------------------------------------------
# the five primitives + ConvNorm (non-slimmable: ratio=None)
# --------------------------------------------------------------------------------------
def conv_norm(x, p: Params, kernel_size=3, stride=1, padding=None, training=False, ratio: Optional[Ratio] = None):
    """ConvNorm.forward, search/operations.py:125-128 (conv -> BN -> ReLU); keys conv.0 / conv.1."""
    if padding is None:
        padding = conv_padding(kernel_size, stride)
    w = p["conv.0.weight"]
    if ratio is not None:
        w = _us_w(w, ratio.ci(w.shape[1]), ratio.co(w.shape[0]))
    y = conv2d(x, w, None, stride, padding)
    y = _us_bn(y, p.sub("conv.1"), None if ratio is None else ratio.idx_out(), training)
    return _q(F.relu(y))


def basic_residual_1x(x, p: Params, stride=1, training=False, ratio: Optional[Ratio] = None):
    """BasicResidual1x.forward, search/operations.py:196-200."""
    w = p["conv1.weight"]
    if ratio is not None:
        w = _us_w(w, ratio.ci(w.shape[1]), ratio.co(w.shape[0]))
    y = conv2d(x, w, None, stride, 1)
    y = _us_bn(y, p.sub("bn1"), None if ratio is None else ratio.idx_out(), training)
    return _q(F.relu(y))


def basic_residual_downup_1x(x, p: Params, stride=1, training=False, ratio: Optional[Ratio] = None):
    """BasicResidual_downup_1x.forward, search/operations.py:270-277."""
    H, W = x.shape[2], x.shape[3]
    y = _q(bilinear_ac(x, (H // 2, W // 2)))
    w = p["conv1.weight"]
    if ratio is not None:
        w = _us_w(w, ratio.ci(w.shape[1]), ratio.co(w.shape[0]))
    y = conv2d(y, w, None, 1, 1)
    y = _us_bn(y, p.sub("bn1"), None if ratio is None else ratio.idx_out(), training)
    if stride == 1:
        y = bilinear_ac(_q(y), (H, W))
    return _q(F.relu(y))


def basic_residual_2x(x, p: Params, stride=1, training=False, ratio: Optional[Ratio] = None):
    """BasicResidual2x.forward, search/operations.py:352-359."""
    w1, w2 = p["conv1.weight"], p["conv2.weight"]
    idx = None
    if ratio is not None:
        co = ratio.co(w1.shape[0])
        w1 = _us_w(w1, ratio.ci(w1.shape[1]), co)
        w2 = _us_w(w2, co, co)  # set_ratio((ratio[1], ratio[1])), operations.py:314
        idx = ratio.idx_out()
    y = _q(F.relu(_us_bn(conv2d(x, w1, None, stride, 1), p.sub("bn1"), idx, training)))
    y = _q(F.relu(_us_bn(conv2d(y, w2, None, 1, 1), p.sub("bn2"), idx, training)))
    return y


def basic_residual_downup_2x(x, p: Params, stride=1, training=False, ratio: Optional[Ratio] = None):
    """BasicResidual_downup_2x.forward, search/operations.py:436-446."""
    H, W = x.shape[2], x.shape[3]
    w1, w2 = p["conv1.weight"], p["conv2.weight"]
    idx = None
    if ratio is not None:
        co = ratio.co(w1.shape[0])
        w1 = _us_w(w1, ratio.ci(w1.shape[1]), co)
        w2 = _us_w(w2, co, co)
        idx = ratio.idx_out()
    y = _q(bilinear_ac(x, (H // 2, W // 2)))
    y = _q(F.relu(_us_bn(conv2d(y, w1, None, 1, 1), p.sub("bn1"), idx, training)))
    y = _us_bn(conv2d(y, w2, None, 1, 1), p.sub("bn2"), idx, training)
    if stride == 1:
        y = bilinear_ac(_q(y), (H, W))
    return _q(F.relu(y))


def factorized_reduce(x, p: Params, stride=1, training=False, ratio: Optional[Ratio] = None, slimmable=None):
    """FactorizedReduce.forward, search/operations.py:521-534.

    stride 2: cat[conv1(x), conv2(x[:, :, 1:, 1:])] -> BN -> ReLU (C_out/2 each, 1x1 stride 2)
    stride 1: slimmable -> 1x1 conv -> BN -> ReLU ; non-slimmable -> identity (no parameters)
    """
    if slimmable is None:
        slimmable = ratio is not None
    if stride == 2:
        w1, w2 = p["conv1.weight"], p["conv2.weight"]
        idx = None
        if ratio is not None:
            # USConv2d(C_in, C_out // 2): out slice = make_divisible((C_out//2) * r_out)
            ci = ratio.ci(w1.shape[1])
            co_half = ratio.co(w1.shape[0])
            w1 = _us_w(w1, ci, co_half)
            w2 = _us_w(w2, ci, co_half)
            idx = ratio.idx_out()
        y = torch.cat([conv2d(x, w1, None, 2, 0), conv2d(x[:, :, 1:, 1:], w2, None, 2, 0)], dim=1)
        return _q(F.relu(_us_bn(y, p.sub("bn"), idx, training)))
    if not slimmable:
        return x
    w = p["conv1.weight"]
    w = _us_w(w, ratio.ci(w.shape[1]), ratio.co(w.shape[0]))
    return _q(F.relu(_us_bn(conv2d(x, w, None, 1, 0), p.sub("bn"), ratio.idx_out(), training)))


OP_FUNCS = [factorized_reduce, basic_residual_1x, basic_residual_downup_1x, basic_residual_2x,
            basic_residual_downup_2x]  # order = PRIMITIVES / OPS_name, search/operations.py:539-552


def conv_bn_relu(x, p: Params, stride=1, pad=0, has_bn=True, has_relu=True, training=False):
    """seg_oprs.ConvBnRelu.forward, search/seg_oprs.py:31-39."""
    y = conv2d(x, p["conv.weight"], p["conv.bias"] if p.has("conv.bias") else None, stride, pad)
    if has_bn:
        y = _bn(y, p.sub("bn"), training)
    return _q(F.relu(y) if has_relu else y)


def feature_fusion(x, p: Params, training=False):
    """FeatureFusion.forward = one 1x1 ConvBnRelu, search/seg_oprs.py:219-222."""
    return conv_bn_relu(x, p.sub("conv_1x1"), 1, 0, training=training)


def head(x, p: Params, training=False):
    """Head.forward: 3x3 ConvBnRelu -> 1x1 conv with bias, search/seg_oprs.py:271-274."""
    y = conv_bn_relu(x, p.sub("conv_3x3"), 1, 1, training=training)
    return _q(conv2d(y, p["conv_1x1.weight"], p["conv_1x1.bias"], 1, 0))


# --------------------------------------------------------------------------------------
# genotype decoder (train/model_seg.py:12-135); pure python/numpy, runs once
# --------------------------------------------------------------------------------------
def _softmax_t(v: torch.Tensor) -> torch.Tensor:
    return F.softmax(v, dim=-1)


def path2downs(path):
    """train/model_seg.py:15-29."""
    downs = [1 if b > a else 0 for a, b in zip(path[:-1], path[1:])]
    for a, b in zip(path[:-1], path[1:]):
        assert (b - a) in (0, 1)
    return downs + [0]


def downs2path(downs):
    """train/model_seg.py:31-38."""
    path = [0]
    for d in downs[:-1]:
        path.append(path[-1] + (1 if d == 1 else 0))
    return path


def betas2path(betas, last, layers):
    """train/model_seg.py:97-114 (betas already softmaxed)."""
    downs = [0] * layers
    if last == 1:
        cand = [float(b[0]) for b in betas[1][1:-1]]
        downs[int(np.argmax(cand)) + 1] = 1
    elif last == 2:
        best, best_ij = 0.0, (0, 1)
        for j in range(layers - 4):
            for i in range(1, j - 1):
                prob = float(betas[1][i][0] * betas[2][j][0])
                if prob > best:
                    best, best_ij = prob, (i, j)
        downs[best_ij[0] + 1] = 1
        downs[best_ij[1] + 2] = 1
    path = downs2path(downs)
    assert path[-1] == last
    return path


def path2widths(path, ratios, width_mult_list):
    """train/model_seg.py:116-124."""
    widths = []
    for layer in range(1, len(path)):
        scale = path[layer]
        row = ratios[scale][layer - 1] if scale == 0 else ratios[scale][layer - scale]
        widths.append(width_mult_list[int(row.argmax())])
    return widths


def alphas2ops_path_width(alphas, path, widths, ignore_skip=False):
    """train/model_seg.py:40-95.  NOTE: mutates `alphas` in place exactly like the reference."""
    assert len(path) == len(widths) + 1
    n = len(path)
    min_len = int(np.round(n / 3.)) + path[-1] * 2
    skips = []  # (position, softmax prob of skip) candidates for pruning
    for i in range(n):
        s = path[i]
        if ignore_skip:
            alphas[s][i - s][0] = -float("inf")
        op = int(alphas[s][i - s].argmax())
        if op == 0 and (i == n - 1 or path[i] == path[i + 1]):
            skips.append((i, _softmax_t(alphas[s][i - s])[0]))
    pos_skips = [pos for pos, _ in skips]
    pos_downs = [pos for pos in range(n - 1) if path[pos] < path[pos + 1]]
    if pos_downs:
        pos_downs.append(n)
        for a, b in zip(pos_downs[:-1], pos_downs[1:]):
            # a whole stretch between two downsamples must not collapse to skips only
            if a + 1 in pos_skips and b - 1 in pos_skips and \
                    pos_skips.index(b - 1) - pos_skips.index(a + 1) == (b - 1) - (a + 1):
                weakest = [1, -1]
                for j in range(a + 1, b):
                    score = _softmax_t(alphas[path[j]][j - path[j]])[0]
                    if score <= weakest[0]:
                        weakest = [score, j]
                j = weakest[1]
                alphas[path[j]][j - path[j]][0] = -float("inf")
    if len(skips) > n - min_len:
        skips = sorted(skips, key=lambda t: t[1], reverse=True)[:n - min_len]
    pos_skips = [pos for pos, _ in skips]
    ops, path_c, widths_c = [], [], []
    for i in range(n):
        s = path[i]
        op = int(alphas[s][i - s].argmax())
        if op == 0:
            if i in pos_skips:
                if i == n - 1:
                    widths_c = widths_c[:-1]
                continue
            alphas[s][i - s][0] = -float("inf")
            op = int(alphas[s][i - s].argmax())
        path_c.append(s)
        if i < len(widths):
            widths_c.append(widths[i])
        ops.append(op)
    assert len(path_c) >= min_len
    return ops, path_c, widths_c


def network_metas(alphas, betas, ratios, width_mult_list, layers, last, ignore_skip=False):
    """train/model_seg.py:126-135.  NOTE: like the reference this re-softmaxes betas[1], betas[2]
    in place on EVERY call (three calls in __init__, :199-201)."""
    betas[1] = _softmax_t(betas[1])
    betas[2] = _softmax_t(betas[2])
    path = betas2path(betas, last, layers)
    widths = path2widths(path, ratios, width_mult_list)
    ops, path, widths = alphas2ops_path_width(alphas, path, widths, ignore_skip=ignore_skip)
    assert len(ops) == len(path) == len(widths) + 1
    return ops, path, path2downs(path), widths


class CellSpec:
    def __init__(self, key, op, c_in, c_out, down, branches):
        self.key, self.op, self.c_in, self.c_out, self.down, self.branches = key, op, c_in, c_out, down, branches

    def as_tuple(self):
        return (self.key, self.op, self.c_in, self.c_out, self.down, list(self.branches))


class StudentStructure:
    """Decoded structure of `Network_Multi_Path_Infer` (train/model_seg.py:174-296)."""

    def __init__(self, alphas, betas, ratios, lasts, layers=16, Fch=12, width_mult_list=WIDTH_MULT_LIST,
                 stem_head_width=(1., 1.), ignore_skip=False, num_classes=19):
        alphas = [a.clone() for a in alphas]
        betas = [None if b is None else b.clone() for b in betas]
        ratios = [r.clone() for r in ratios]
        self.Fch, self.num_classes, self.layers = Fch, num_classes, layers
        if ratios[0].size(1) == 1:  # model_seg.py:183-189
            width_mult_list = [1.] if ignore_skip else [4. / 12]
        self.width_mult_list = list(width_mult_list)
        self.stem_head_width = tuple(stem_head_width)
        self.metas = {}
        for last in (0, 1, 2):  # model_seg.py:199-201 (order matters: betas mutate)
            self.metas[last] = network_metas(alphas, betas, ratios, self.width_mult_list, layers, last, ignore_skip)
        self.lasts = list(lasts)
        self.branch = len(lasts)
        self.ops = [self.metas[l][0] for l in lasts]
        self.paths = [self.metas[l][1] for l in lasts]
        self.downs = [self.metas[l][2] for l in lasts]
        self.widths = [self.metas[l][3] for l in lasts]
        self._group_cells()

    def nf(self, scale, width=1.0):
        return num_filters(scale, self.Fch, width)

    def _group_cells(self):
        """get_branch_groups_cells, train/model_seg.py:241-296."""
        ops, paths, downs, widths, lasts = self.ops, self.paths, self.downs, self.widths, self.lasts
        nb = self.branch
        n_layers = max(len(p) for p in paths)
        self.ch_16 = self.ch_8_2 = self.ch_8_1 = 0
        self.branch_groups: List[List[List[int]]] = []
        self.cells: Dict[str, CellSpec] = {}
        connected = np.ones((nb, nb))
        for l in range(n_layers):
            conn = np.ones((nb, nb))
            for i in range(nb):
                for j in range(i + 1, nb):
                    if len(paths[i]) <= l + 1 or len(paths[j]) <= l + 1 or paths[i][l + 1] != paths[j][l + 1] \
                            or ops[i][l] != ops[j][l] or widths[i][l] != widths[j][l]:
                        conn[i, j] = conn[j, i] = 0
            connected *= conn
            groups: List[List[int]] = []
            for b in range(nb):
                if len(paths[b]) < l + 1:
                    continue
                placed = False
                for g in groups:
                    if connected[g[0], b] == 1:
                        g.append(b)
                        placed = True
                if not placed:
                    groups.append([b])
            for g in groups:
                b0 = g[0]
                op = ops[b0][l]
                scale = 2 ** (paths[b0][l] + 3)
                down = downs[b0][l]
                shw = self.stem_head_width
                if l == 0:
                    c_in, c_out = self.nf(scale, shw[0]), self.nf(scale * (down + 1), widths[b0][l])
                elif l == len(paths[b0]) - 1:
                    assert down == 0
                    c_in, c_out = self.nf(scale, widths[b0][l - 1]), self.nf(scale, shw[1])
                else:
                    c_in, c_out = self.nf(scale, widths[b0][l - 1]), self.nf(scale * (down + 1), widths[b0][l])
                if 2 in lasts and lasts.index(2) in g and down and scale == 16:
                    self.ch_16 = c_in
                if 2 in lasts and lasts.index(2) in g and down and scale == 8:
                    self.ch_8_2 = c_in
                if 1 in lasts and lasts.index(1) in g and down and scale == 8:
                    self.ch_8_1 = c_in
                spec = CellSpec("%d-%d" % (l, b0), int(op), c_in, c_out, int(down), list(g))
                for b in g:
                    self.cells["%d-%d" % (l, b)] = spec
            self.branch_groups.append(groups)

    def describe(self):
        return {
            "lasts": self.lasts,
            "ops": [[int(o) for o in ops] for ops in self.ops],
            "paths": [[int(v) for v in p] for p in self.paths],
            "downs": [[int(v) for v in d] for d in self.downs],
            "widths": [[float(v) for v in w] for w in self.widths],
            "branch_groups": self.branch_groups,
            "ch_16": self.ch_16, "ch_8_2": self.ch_8_2, "ch_8_1": self.ch_8_1,
            "cells": {k: v.as_tuple() for k, v in self.cells.items()},
        }


def student_forward(x, sd: Dict[str, torch.Tensor], st: StudentStructure, training=False,
                    return_pred8_lowres=False):
    """Network_Multi_Path_Infer.forward + agg_ffm, train/model_seg.py:298-366.

    eval : returns logits upsampled x8 (bilinear, align_corners=True), shape (N,19,H,W)
    train: returns (pred8, pred16, pred32) each upsampled to full resolution (:357-362)
    """
    P = Params(sd)
    H = x.shape[2]
    y = conv_norm(x, P.sub("stem.0"), 3, 2, 1, training)
    y = basic_residual_2x(y, P.sub("stem.1"), 2, training)
    y = basic_residual_2x(y, P.sub("stem.2"), 2, training)
    nb = st.branch
    outputs = [y] * nb
    out8, out16, out32 = [y] * nb, [y] * nb, [y] * nb
    for layer, groups in enumerate(st.branch_groups):
        for g in groups:
            spec = st.cells["%d-%d" % (layer, g[0])]
            p = P.sub("cells.%d-%d._op._op" % (layer, g[0]))
            stride = 2 if spec.down else 1
            o = OP_FUNCS[spec.op](outputs[g[0]], p, stride, training)
            scale = int(H // o.shape[2])
            for b in g:
                outputs[b] = o
                if scale == 8:
                    out8[b] = o
                elif scale == 16:
                    out16[b] = o
                elif scale == 32:
                    out32[b] = o
    pred32, pred16, pred8 = [], [], []
    for b in range(nb):
        last = st.lasts[b]
        if last == 2:
            if training:
                pred32.append(out32[b])
            o = conv_norm(out32[b], P.sub("arms32.0"), 1, 1, 0, training)
            o = _q(bilinear_ac(o, out16[b].shape[2:]))
            o = conv_norm(torch.cat([o, out16[b]], 1), P.sub("refines32.0"), 3, 1, 1, training)
            if training:
                pred16.append(out16[b])
            o = conv_norm(o, P.sub("arms32.1"), 1, 1, 0, training)
            o = _q(bilinear_ac(o, out8[b].shape[2:]))
            o = conv_norm(torch.cat([o, out8[b]], 1), P.sub("refines32.1"), 3, 1, 1, training)
            pred8.append(o)
        elif last == 1:
            if training:
                pred16.append(out16[b])
            o = conv_norm(out16[b], P.sub("arms16"), 1, 1, 0, training)
            o = _q(bilinear_ac(o, out8[b].shape[2:]))
            o = conv_norm(torch.cat([o, out8[b]], 1), P.sub("refines16"), 3, 1, 1, training)
            pred8.append(o)
        else:
            pred8.append(out8[b])
    p8 = head(feature_fusion(torch.cat(pred8, 1), P.sub("ffm"), training), P.sub("heads8"), training)
    if not training:
        if return_pred8_lowres:
            return p8
        return bilinear_ac(p8, (p8.shape[2] * 8, p8.shape[3] * 8))
    p16 = head(torch.cat(pred16, 1), P.sub("heads16"), training) if pred16 else None
    p32 = head(torch.cat(pred32, 1), P.sub("heads32"), training) if pred32 else None
    p8 = bilinear_ac(p8, (p8.shape[2] * 8, p8.shape[3] * 8))
    if p16 is not None:
        p16 = bilinear_ac(p16, (p16.shape[2] * 16, p16.shape[3] * 16))
    if p32 is not None:
        p32 = bilinear_ac(p32, (p32.shape[2] * 32, p32.shape[3] * 32))
    return p8, p16, p32


def student_param_shapes(st: StudentStructure, training=False) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape for the conv/BN tensors the forward uses (eval build omits heads16/32,
    train/model_seg.py:217-226).  Used to synthesise random weights without the reference."""
    shapes: Dict[str, Tuple[int, ...]] = {}

    def bn(prefix, c):
        shapes[prefix + ".weight"] = (c,)
        shapes[prefix + ".bias"] = (c,)
        shapes[prefix + ".running_mean"] = (c,)
        shapes[prefix + ".running_var"] = (c,)

    def convnorm(prefix, ci, co, k):
        shapes[prefix + ".conv.0.weight"] = (co, ci, k, k)
        bn(prefix + ".conv.1", co)

    def op(prefix, op_idx, ci, co, stride):
        if op_idx in (1, 2):
            shapes[prefix + ".conv1.weight"] = (co, ci, 3, 3)
            bn(prefix + ".bn1", co)
        elif op_idx in (3, 4):
            shapes[prefix + ".conv1.weight"] = (co, ci, 3, 3)
            bn(prefix + ".bn1", co)
            shapes[prefix + ".conv2.weight"] = (co, co, 3, 3)
            bn(prefix + ".bn2", co)
        elif op_idx == 0 and stride == 2:
            shapes[prefix + ".conv1.weight"] = (co // 2, ci, 1, 1)
            shapes[prefix + ".conv2.weight"] = (co // 2, ci, 1, 1)
            bn(prefix + ".bn", co)

    def headp(prefix, cin):
        mid = cin if cin <= 256 else cin // 2  # seg_oprs.py:231-244
        shapes[prefix + ".conv_3x3.conv.weight"] = (mid, cin, 3, 3)
        bn(prefix + ".conv_3x3.bn", mid)
        shapes[prefix + ".conv_1x1.weight"] = (st.num_classes, mid, 1, 1)
        shapes[prefix + ".conv_1x1.bias"] = (st.num_classes,)

    shw = st.stem_head_width
    c2, c4, c8 = st.nf(2, shw[0]) * 2, st.nf(4, shw[0]) * 2, st.nf(8, shw[0])
    convnorm("stem.0", 3, c2, 3)
    op("stem.1", 3, c2, c4, 2)
    op("stem.2", 3, c4, c8, 2)
    for key, spec in st.cells.items():
        op("cells.%s._op._op" % key, spec.op, spec.c_in, spec.c_out, 2 if spec.down else 1)
    h8, h16, h32 = st.nf(8, shw[1]), st.nf(16, shw[1]), st.nf(32, shw[1])
    if 2 in st.lasts:
        convnorm("arms32.0", h32, h16, 1)
        convnorm("arms32.1", h16, h8, 1)
        convnorm("refines32.0", h16 + st.ch_16, h16, 3)
        convnorm("refines32.1", h8 + st.ch_8_2, h8, 3)
    if 1 in st.lasts:
        convnorm("arms16", h16, h8, 1)
        convnorm("refines16", h8 + st.ch_8_1, h8, 3)
    shapes["ffm.conv_1x1.conv.weight"] = (h8 * st.branch, h8 * st.branch, 1, 1)
    bn("ffm.conv_1x1.bn", h8 * st.branch)
    headp("heads8", h8 * st.branch)
    if training:
        if 2 in st.lasts:
            headp("heads32", h32)
            headp("heads16", h16 + st.ch_16 if 1 in st.lasts else st.ch_16)
        else:
            headp("heads16", h16)
    return shapes


def random_state_dict(shapes: Dict[str, Tuple[int, ...]], seed=12345, randomize_bn=True):
    """Synthetic weights per SURVEY.md section 8(d): kaiming_normal(fan_in, relu) convs; BN gamma/beta and
    running stats randomised so that eval-mode BN is not a no-op.  numpy's legacy MT19937
    (`RandomState`) is used because its stream is stable across platforms and versions, so the
    build container and the GPU box synthesise bit-identical weights from the same seed."""
    rs = np.random.RandomState(seed)
    sd = {}

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    for k in sorted(shapes):
        shp = shapes[k]
        if len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = t(rs.standard_normal(shp) * math.sqrt(2.0 / fan_in))
        elif k.endswith("running_var"):
            sd[k] = t(rs.uniform(0.5, 1.5, shp)) if randomize_bn else torch.ones(shp)
        elif k.endswith("running_mean"):
            sd[k] = t(rs.standard_normal(shp) * 0.1) if randomize_bn else torch.zeros(shp)
        elif k.endswith("conv_1x1.bias"):  # Head's 1x1 conv bias, seg_oprs.py:246
            sd[k] = t(rs.standard_normal(shp) * 0.05)
        elif k.endswith(".weight"):  # BN gamma
            sd[k] = t(1.0 + 0.1 * rs.standard_normal(shp)) if randomize_bn else torch.ones(shp)
        else:  # BN beta
            sd[k] = t(0.1 * rs.standard_normal(shp)) if randomize_bn else torch.zeros(shp)
    return sd


def random_input(shape, seed=12345):
    """Synthetic image batch ~ N(0,1) (the reference times with torch.randn, darts_utils.py:189)."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal(shape).astype(np.float32))


# --------------------------------------------------------------------------------------
# losses used by callers' parity tests (tools/seg_opr/loss_opr.py:43-93) -- "next" row N1
# --------------------------------------------------------------------------------------
def ohem_cross_entropy(pred, target, ignore_label=255, thresh=0.7, min_kept=256):
    """ProbOhemCrossEntropy2d.forward, tools/seg_opr/loss_opr.py:63-93."""
    b, c, h, w = pred.shape
    t = target.reshape(-1).clone()
    valid = t.ne(ignore_label)
    t = t * valid.long()
    num_valid = int(valid.sum())
    prob = F.softmax(pred, dim=1).transpose(0, 1).reshape(c, -1)
    if min_kept > num_valid:
        pass
    elif num_valid > 0:
        prob = prob.masked_fill(~valid, 1)
        mask_prob = prob[t, torch.arange(len(t))]
        threshold = thresh
        if min_kept > 0:
            index = mask_prob.argsort()
            kth = index[min(len(index), min_kept) - 1]
            if mask_prob[kth] > thresh:
                threshold = mask_prob[kth]
            kept = mask_prob.le(threshold)
            t = t * kept.long()
            valid = valid & kept
    t = t.masked_fill(~valid, ignore_label).view(b, h, w)
    return F.cross_entropy(pred, t, ignore_index=ignore_label)


def distill_kl(student_logits, teacher_logits):
    """Distillation term of train/train.py:64,260: nn.KLDivLoss() (reduction 'mean' = mean over ALL elements) of
    log(softmax(student)) against softmax(teacher), both over the class dimension."""
    log_p = F.softmax(student_logits, dim=1).log()
    q = F.softmax(teacher_logits, dim=1)
    return (q * (q.log() - log_p)).mean()

