"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the search supernet (search/model_search.py:46-505).

Functional like oracle/fasterseg_oracle.py: parameters (weights, BN buffers AND the architecture parameters
alpha_*/beta_*/ratio_*) come from a dict keyed by the reference's state_dict names.  Pinned to golden vectors produced by the
unmodified reference in tests/golden/supernet.npz (oracle/make_golden_supernet.py).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import fasterseg_oracle as orc
from .fasterseg_oracle import Params, Ratio, _q


class SupernetConfig:
    def __init__(self, layers=16, Fch=12, width_mult_list=orc.WIDTH_MULT_LIST, prun_modes=("max", "arch_ratio"),
                 stem_head_width=((1., 1.), (8. / 12, 8. / 12)), num_classes=19):
        self.layers, self.Fch, self.wml = layers, Fch, list(width_mult_list)
        self.prun_modes, self.shw, self.num_classes = list(prun_modes), [tuple(s) for s in stem_head_width], num_classes

    def nf(self, scale, width=1.0):
        return orc.num_filters(scale, self.Fch, width)


def sample_gumbel(shape, eps=1e-20):
    """model_search.py:14-17 (CPU generator stream)."""
    U = torch.rand(shape)
    return -torch.log(-torch.log(U + eps) + eps)


def gumbel_softmax_hard(logits):
    """model_search.py:20-43 with hard=True."""
    y = F.softmax(logits + sample_gumbel(logits.size()), dim=-1)
    ind = y.argmax(dim=-1)
    y_hard = torch.zeros_like(y)
    y_hard[ind] = 1.0
    return (y_hard - y).detach() + y


def sample_prun_ratio(sd, cfg: SupernetConfig, arch_idx: int, mode: str):
    """model_search.py:209-261; the numpy / torch global RNG streams are consumed in the reference's order."""
    counts = (cfg.layers - 1, cfg.layers - 1, cfg.layers - 2)
    if mode == "arch_ratio":
        out = []
        for s, n in zip(range(3), counts):
            param = sd["ratio_%d_%d" % (arch_idx, s)]
            out.append([gumbel_softmax_hard(F.log_softmax(param[layer], dim=-1)) for layer in range(n)])
        return out
    if mode == "min":
        return [[cfg.wml[0]] * n for n in counts]
    if mode == "max":
        return [[cfg.wml[-1]] * n for n in counts]
    return [[np.random.choice(cfg.wml) for _ in range(n)] for n in counts]


def _resolve(r, wml):
    if isinstance(r, torch.Tensor):
        i = int(r.argmax())
        return wml[i], r[i]
    return r, 1.0


def mixed_op(x, P: Params, weights, ratios, stride, cfg: SupernetConfig, training: bool):
    """MixedOp.forward, model_search.py:60-78: sum_k op_k(x) * w_k * r0 * r1."""
    ratio0, s0 = _resolve(ratios[0], cfg.wml)
    ratio1, s1 = _resolve(ratios[1], cfg.wml)
    ratio = Ratio(ratio0, ratio1, cfg.wml)
    result = 0
    for k, fn in enumerate(orc.OP_FUNCS):
        p = P.sub("_ops.%d" % k)
        if k == 0:
            y = orc.factorized_reduce(x, p, stride, training, ratio, slimmable=True)
        else:
            y = fn(x, p, stride, training, ratio)
        result = result + y * weights[k] * s0 * s1
    return _q(result)


def cell(x, P: Params, alphas, ratios, down: bool, cfg, training):
    """Cell.forward, model_search.py:116-121."""
    out = mixed_op(x, P.sub("_op"), alphas, (ratios[0], ratios[1]), 1, cfg, training)
    dn = mixed_op(x, P.sub("downsample"), alphas, (ratios[0], ratios[2]), 2, cfg, training) if down else None
    return out, dn


def _ratio_triple(i, j, ratios, cfg, arch_idx):
    shw = cfg.shw[arch_idx]
    L = cfg.layers
    if i == 0 and j == 0:
        return (shw[0], ratios[j][i - j], ratios[j + 1][i - j])
    if i == L - 1:
        return (ratios[j][i - j - 1] if j == 0 else ratios[j][i - j], shw[1], None)
    if j == 2:
        return (ratios[j][i - j], ratios[j][i - j + 1], None)
    if j == 0:
        return (ratios[j][i - j - 1], ratios[j][i - j], ratios[j + 1][i - j])
    return (ratios[j][i - j], ratios[j][i - j + 1], ratios[j + 1][i - j])


def _cell_has_down(l, j, L):
    if l == 0 or l == 1:
        return True
    if l < L - 1:
        return j < 2
    return False


def supernet_forward(x, sd: Dict[str, torch.Tensor], cfg: SupernetConfig, arch_idx=0, prun_mode=None, training=True):
    """Network_Multi_Path.forward, model_search.py:263-358 -> (pred0, pred1, pred2, pred02, pred12)."""
    P = Params(sd)
    alphas = [F.softmax(sd["alpha_%d_%d" % (arch_idx, s)], dim=-1) for s in range(3)]
    betas = [None, F.softmax(sd["beta_%d_1" % arch_idx], dim=-1), F.softmax(sd["beta_%d_2" % arch_idx], dim=-1)]
    mode = prun_mode if prun_mode is not None else cfg.prun_modes[arch_idx]
    ratios = sample_prun_ratio(sd, cfg, arch_idx, mode)
    L = cfg.layers

    ps = P.sub("stem.%d" % arch_idx)
    y = orc.conv_norm(x, ps.sub("0"), 3, 2, 1, training)
    y = orc.basic_residual_2x(y, ps.sub("1"), 2, training)
    y = orc.basic_residual_2x(y, ps.sub("2"), 2, training)
    out_prev = [[y, None]]
    for i in range(L):
        n_scales = 1 if i == 0 else (2 if i == 1 else 3)
        out = []
        for j in range(n_scales):
            pc = P.sub("cells.%d.%d" % (i, j))
            alpha = alphas[j][i - j]
            ratio = _ratio_triple(i, j, ratios, cfg, arch_idx)
            down = _cell_has_down(i, j, L)
            if j == 0:
                out.append(cell(out_prev[0][0], pc, alpha, ratio, down, cfg, training))
            elif i == j:
                out.append(cell(out_prev[j - 1][1], pc, alpha, ratio, down, cfg, training))
            else:
                b = betas[j][i - j - 1]
                out0 = down0 = out1 = down1 = None
                if b[0] > 0:
                    out0, down0 = cell(out_prev[j - 1][1], pc, alpha, ratio, down, cfg, training)
                if b[1] > 0:
                    out1, down1 = cell(out_prev[j][0], pc, alpha, ratio, down, cfg, training)
                o = _q(sum(w * t for w, t in zip(b, [out0, out1])))
                d = sum(w * t if t is not None else 0 for w, t in zip(b, [down0, down1]))
                out.append((o, _q(d) if torch.is_tensor(d) else d))
        out_prev = out

    def up2(t):
        return _q(orc.bilinear_ac(t, (t.shape[2] * 2, t.shape[3] * 2)))

    r16, r32 = P.sub("refine16.%d" % arch_idx), P.sub("refine32.%d" % arch_idx)
    out0 = out[0][0]
    out1 = orc.conv_norm(torch.cat([up2(orc.conv_norm(out[1][0], r16.sub("0"), 1, 1, None, training)), out[0][0]], 1),
                         r16.sub("1"), 3, 1, 1, training)
    out2 = orc.conv_norm(torch.cat([up2(orc.conv_norm(out[2][0], r32.sub("0"), 1, 1, None, training)), out[1][0]], 1),
                         r32.sub("1"), 3, 1, 1, training)
    out2 = orc.conv_norm(torch.cat([up2(orc.conv_norm(out2, r32.sub("2"), 1, 1, None, training)), out[0][0]], 1),
                         r32.sub("3"), 3, 1, 1, training)
    preds = [orc.head(out0, P.sub("head0.%d" % arch_idx), training), orc.head(out1, P.sub("head1.%d" % arch_idx), training),
             orc.head(out2, P.sub("head2.%d" % arch_idx), training),
             orc.head(torch.cat([out0, out2], 1), P.sub("head02.%d" % arch_idx), training),
             orc.head(torch.cat([out1, out2], 1), P.sub("head12.%d" % arch_idx), training)]
    if not training:
        preds = [orc.bilinear_ac(p, (p.shape[2] * 8, p.shape[3] * 8)) for p in preds]
    return tuple(preds)


def supernet_loss(x, target, sd, cfg: SupernetConfig, criterion, pretrain=False, training=True):
    """Network_Multi_Path._loss, model_search.py:478-505."""
    loss = 0
    if pretrain is not True:
        for idx in range(len(cfg.prun_modes)):
            logits = supernet_forward(x, sd, cfg, idx, None, training)
            loss = loss + sum(criterion(l, target) for l in logits)
        arch_idx = len(cfg.prun_modes) - 1  # the reference leaves self.arch_idx at the last architecture
    else:
        arch_idx = 0
    if len(cfg.wml) > 1:
        modes = ["max", "min"] + (["random", "random"] if pretrain is True else [])
        for mode in modes:
            logits = supernet_forward(x, sd, cfg, arch_idx, mode, training)
            loss = loss + sum(criterion(l, target) for l in logits)
    elif pretrain is True:
        logits = supernet_forward(x, sd, cfg, arch_idx, "max", training)
        loss = loss + sum(criterion(l, target) for l in logits)
    return loss
