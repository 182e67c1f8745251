"""Multi-resolution search supernet -- drop-in for the reference's search/model_search.py.

Kept (search/model_search.py:14-548): gumbel helpers, `MixedOp` / `Cell` / `Network_Multi_Path` with the same constructor
arguments, `sample_prun_ratio`, `forward`, `forward_latency`, `_loss`, `_build_arch_parameters`, `_reset_arch_parameters`,
attributes (`_arch_names`, `_arch_parameters`, `arch_idx`, `prun_mode`, `_prun_modes` ...) and every parameter name
(`stem.0.0.conv.0.weight`, `cells.3.1._op._ops.4.bn2.bn.2.running_var`, `alpha_0_1`, `beta_1_2`, `ratio_1_0` ...), so
search/train_search.py and search/architect.py drive it unchanged.

B200 side: NHWC fp16 activations; every conv+BN+ReLU is one fused tcgen05 kernel (train mode: conv with fused statistics,
finalize, apply); the `result + op(x) * w * r0 * r1` accumulation over the five primitives and the beta-weighted mix of
the two cell invocations are single weighted-sum kernels (K5) whose backward also yields the scalar gradients of
alphas / betas / ratios; logits leave as NCHW fp32 through one layout kernel.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Variable

from . import autograd as AG
from . import functional as F_
from .genotypes import PRIMITIVES
from .operations import *  # noqa: F401,F403
from .operations import OPS, BasicResidual2x, ConvNorm
from .seg_oprs import Head


# https://github.com/YongfeiYan/Gumbel_Softmax_VAE (as cited by the reference, model_search.py:13)
def sample_gumbel(shape, eps=1e-20, device=None):
    U = torch.rand(shape)  # CPU generator, like the reference (which then moves U to the GPU)
    if device is not None:
        U = U.to(device)
    return -torch.log(-torch.log(U + eps) + eps)


def gumbel_softmax_sample(logits, temperature=1):
    y = logits + sample_gumbel(logits.size(), device=logits.device)
    return F.softmax(y / temperature, dim=-1)


def gumbel_softmax(logits, temperature=1, hard=False):
    """ST-gumbel-softmax: one-hot forward value, soft gradient."""
    y = gumbel_softmax_sample(logits, temperature)
    if not hard:
        return y
    shape = y.size()
    _, ind = y.max(dim=-1)
    y_hard = torch.zeros_like(y).view(-1, shape[-1])
    y_hard.scatter_(1, ind.view(-1, 1), 1)
    y_hard = y_hard.view(*shape)
    return (y_hard - y).detach() + y


def _resolve_ratio(r, width_mult_list):
    """tensor -> (width_mult_list[argmax], score tensor = r[argmax]) ; float -> (forced width, 1.)  (model_search.py:63-74)"""
    if isinstance(r, torch.Tensor):
        idx = int(r.argmax())
        return width_mult_list[idx], r[idx]
    return r, 1.


class MixedOp(nn.Module):
    def __init__(self, C_in, C_out, stride=1, width_mult_list=[1.]):
        super(MixedOp, self).__init__()
        self._ops = nn.ModuleList()
        self._width_mult_list = width_mult_list
        for primitive in PRIMITIVES:
            self._ops.append(OPS[primitive](C_in, C_out, stride, True, width_mult_list=width_mult_list))

    def set_prun_ratio(self, ratio):
        for op in self._ops:
            op.set_ratio(ratio)

    def _scaled_weights(self, weights, ratios):
        ratio0, r_score0 = _resolve_ratio(ratios[0], self._width_mult_list)
        ratio1, r_score1 = _resolve_ratio(ratios[1], self._width_mult_list)
        self.set_prun_ratio((ratio0, ratio1))
        return weights * r_score0 * r_score1  # [len(PRIMITIVES)] scalar arithmetic (plumbing)

    def forward(self, x, weights, ratios):
        # int: force #channel; tensor: arch_ratio; float(<=1): force width
        wvec = self._scaled_weights(weights, ratios)
        outs = [F_.to_nhwc_half(op(x)) for op in self._ops]
        if torch.is_grad_enabled() and (wvec.requires_grad or any(o.requires_grad for o in outs)):
            return AG.weighted_sum(wvec, outs)
        return F_.wsum_fwd(outs, wvec.detach().float().contiguous())

    def forward_latency(self, size, weights, ratios):
        wvec = self._scaled_weights(weights, ratios)
        result = 0
        for w, op in zip(wvec, self._ops):
            latency, size_out = op.forward_latency(size)
            result = result + latency * w
        return result, size_out


class Cell(nn.Module):
    def __init__(self, C_in, C_out=None, down=True, width_mult_list=[1.]):
        super(Cell, self).__init__()
        self._C_in = C_in
        if C_out is None: C_out = C_in
        self._C_out = C_out
        self._down = down
        self._width_mult_list = width_mult_list
        self._op = MixedOp(C_in, C_out, width_mult_list=width_mult_list)
        if self._down:
            self.downsample = MixedOp(C_in, C_in * 2, stride=2, width_mult_list=width_mult_list)

    def forward(self, input, alphas, ratios):
        # ratios: (in, out, down)
        out = self._op(input, alphas, (ratios[0], ratios[1]))
        assert (self._down and (ratios[2] is not None)) or ((not self._down) and (ratios[2] is None))
        down = self.downsample(input, alphas, (ratios[0], ratios[2])) if self._down else None
        return out, down

    def forward_latency(self, size, alphas, ratios):
        out = self._op.forward_latency(size, alphas, (ratios[0], ratios[1]))
        assert (self._down and (ratios[2] is not None)) or ((not self._down) and (ratios[2] is None))
        down = self.downsample.forward_latency(size, alphas, (ratios[0], ratios[2])) if self._down else None
        return out, down


def _mix(betas_row, a, b):
    """beta-weighted sum of the two invocations of a cell (model_search.py:330-333)."""
    if a is None and b is None:
        return 0
    a, b = F_.to_nhwc_half(a), F_.to_nhwc_half(b)
    if torch.is_grad_enabled() and (betas_row.requires_grad or a.requires_grad or b.requires_grad):
        return AG.weighted_sum(betas_row, [a, b])
    return F_.wsum_fwd([a, b], betas_row.detach().float().contiguous())


class Network_Multi_Path(nn.Module):
    def __init__(self, num_classes=19, layers=16, criterion=nn.CrossEntropyLoss(ignore_index=-1), Fch=12, width_mult_list=[1., ],
                 prun_modes=['arch_ratio', ], stem_head_width=[(1., 1.), ]):
        super(Network_Multi_Path, self).__init__()
        self._num_classes = num_classes
        assert layers >= 3
        self._layers = layers
        self._criterion = criterion
        self._Fch = Fch
        self._width_mult_list = width_mult_list
        self._prun_modes = prun_modes
        self.prun_mode = None  # prun_mode is higher priority than _prun_modes
        self._stem_head_width = stem_head_width
        self._flops = 0
        self._params = 0

        nf = self.num_filters
        self.stem = nn.ModuleList([
            nn.Sequential(
                ConvNorm(3, nf(2, sr) * 2, kernel_size=3, stride=2, padding=1, bias=False, groups=1, slimmable=False),
                BasicResidual2x(nf(2, sr) * 2, nf(4, sr) * 2, kernel_size=3, stride=2, groups=1, slimmable=False),
                BasicResidual2x(nf(4, sr) * 2, nf(8, sr), kernel_size=3, stride=2, groups=1, slimmable=False))
            for sr, _ in self._stem_head_width])

        self.cells = nn.ModuleList()
        for l in range(layers):
            if l == 0:
                scales, downs = [8], [True]
            elif l == 1:
                scales, downs = [8, 16], [True, True]
            elif l < layers - 1:
                scales, downs = [8, 16, 32], [True, True, False]
            else:
                scales, downs = [8, 16, 32], [False, False, False]
            self.cells.append(nn.ModuleList(Cell(nf(s), down=d, width_mult_list=width_mult_list) for s, d in zip(scales, downs)))

        def cn(ci, co, k):
            return ConvNorm(ci, co, kernel_size=k, padding=1 if k == 3 else None, bias=False, groups=1, slimmable=False)

        self.refine32 = nn.ModuleList([
            nn.ModuleList([cn(nf(32, hr), nf(16, hr), 1), cn(nf(32, hr), nf(16, hr), 3), cn(nf(16, hr), nf(8, hr), 1),
                           cn(nf(16, hr), nf(8, hr), 3)]) for _, hr in self._stem_head_width])
        self.refine16 = nn.ModuleList([
            nn.ModuleList([cn(nf(16, hr), nf(8, hr), 1), cn(nf(16, hr), nf(8, hr), 3)]) for _, hr in self._stem_head_width])
        self.head0 = nn.ModuleList([Head(nf(8, hr), num_classes, False) for _, hr in self._stem_head_width])
        self.head1 = nn.ModuleList([Head(nf(8, hr), num_classes, False) for _, hr in self._stem_head_width])
        self.head2 = nn.ModuleList([Head(nf(8, hr), num_classes, False) for _, hr in self._stem_head_width])
        self.head02 = nn.ModuleList([Head(nf(8, hr) * 2, num_classes, False) for _, hr in self._stem_head_width])
        self.head12 = nn.ModuleList([Head(nf(8, hr) * 2, num_classes, False) for _, hr in self._stem_head_width])

        # arch parameter names: {"alphas": [...], "betas": [...], "ratios": [...]} per architecture (teacher / student)
        self._arch_names = []
        self._arch_parameters = []
        for i in range(len(self._prun_modes)):
            arch_name, arch_param = self._build_arch_parameters(i)
            self._arch_names.append(arch_name)
            self._arch_parameters.append(arch_param)
            self._reset_arch_parameters(i)
        self.arch_idx = 0  # which architecture's stem / heads / arch parameters the next forward uses

    def num_filters(self, scale, width=1.0):
        return int(np.round(scale * self._Fch * width))

    def new(self):
        """(dead in the reference too: it references an undefined `Network`, model_search.py:203-207)"""
        model_new = Network_Multi_Path(self._num_classes, self._layers, self._criterion, self._Fch).cuda()
        for x, y in zip(model_new._arch_parameters, self._arch_parameters):
            for a, b in zip(x, y):
                a.data.copy_(b.data)
        return model_new

    def sample_prun_ratio(self, mode="arch_ratio"):
        '''mode: "min"|"max"|"random"|"arch_ratio"(default)'''
        assert mode in ["min", "max", "random", "arch_ratio"]
        counts = (self._layers - 1, self._layers - 1, self._layers - 2)
        if mode == "arch_ratio":
            names = self._arch_names[self.arch_idx]["ratios"]
            out = []
            for name, n in zip(names, counts):
                param = getattr(self, name)
                out.append([gumbel_softmax(F.log_softmax(param[layer], dim=-1), hard=True) for layer in range(n)])
            return out
        if mode == "min":
            pick = lambda: self._width_mult_list[0]
        elif mode == "max":
            pick = lambda: self._width_mult_list[-1]
        else:
            pick = lambda: np.random.choice(self._width_mult_list)
        # sampling order (scale 0 layers, then scale 1, then scale 2) matters for the shared numpy RNG stream
        return [[pick() for _ in range(n)] for n in counts]

    def _arch(self, kind, i):
        return getattr(self, self._arch_names[self.arch_idx][kind][i])

    def _ratio_triple(self, i, j, ratios):
        """(in, out, down) width ratios of the cell at layer i, scale j (model_search.py:300-316)."""
        shw = self._stem_head_width[self.arch_idx]
        if i == 0 and j == 0:
            return (shw[0], ratios[j][i - j], ratios[j + 1][i - j])
        if i == self._layers - 1:
            return (ratios[j][i - j - 1] if j == 0 else ratios[j][i - j], shw[1], None)
        if j == 2:
            return (ratios[j][i - j], ratios[j][i - j + 1], None)
        if j == 0:
            return (ratios[j][i - j - 1], ratios[j][i - j], ratios[j + 1][i - j])
        return (ratios[j][i - j], ratios[j][i - j + 1], ratios[j + 1][i - j])

    def forward(self, input):
        # out_prev: cell-state; index 0: keep; index 1: down
        idx = self.arch_idx
        stem, refine16, refine32 = self.stem[idx], self.refine16[idx], self.refine32[idx]
        alphas = [F.softmax(self._arch("alphas", s), dim=-1) for s in range(3)]
        betas = [None, F.softmax(self._arch("betas", 0), dim=-1), F.softmax(self._arch("betas", 1), dim=-1)]
        # one host read instead of a GPU->CPU sync per `betas[...] > 0` test (model_search.py:326-329)
        betas_pos = [None] + [(b.detach() > 0).tolist() for b in betas[1:]]
        ratios = self.sample_prun_ratio(mode=self.prun_mode if self.prun_mode is not None else self._prun_modes[idx])

        out_prev = [[stem(input), None]]  # stem: one cell
        for i, cells in enumerate(self.cells):      # i: layer
            out = []
            for j, cell in enumerate(cells):        # j: scale
                alpha = alphas[j][i - j]
                ratio = self._ratio_triple(i, j, ratios)
                if j == 0:
                    out.append(cell(out_prev[0][0], alpha, ratio))
                elif i == j:
                    out.append(cell(out_prev[j - 1][1], alpha, ratio))
                else:
                    # the cell runs twice with the same weights: on the downsampled output of the scale above ("0: from
                    # down") and on its own previous output ("1: from keep"); BN running stats see both, in this order
                    out0 = down0 = out1 = down1 = None
                    if betas_pos[j][i - j - 1][0]:
                        out0, down0 = cell(out_prev[j - 1][1], alpha, ratio)
                    if betas_pos[j][i - j - 1][1]:
                        out1, down1 = cell(out_prev[j][0], alpha, ratio)
                    brow = betas[j][i - j - 1]
                    if out0 is None or out1 is None:  # a beta underflowed to 0: plain scaled term, like the reference's sum()
                        keep = out1 if out0 is None else out0
                        keepd = down1 if out0 is None else down0
                        w = brow[1] if out0 is None else brow[0]
                        out.append((_mix(torch.stack([w, w * 0]), keep, keep), 0 if keepd is None else _mix(torch.stack([w, w * 0]), keepd, keepd)))
                    else:
                        out.append((_mix(brow, out0, out1), _mix(brow, down0, down1) if down0 is not None else 0))
            out_prev = out

        up2 = lambda t: _resize2x(t)
        out0 = out[0][0]
        out1 = refine16[1](_cat([up2(refine16[0](out[1][0])), out[0][0]]))
        out2 = refine32[1](_cat([up2(refine32[0](out[2][0])), out[1][0]]))
        out2 = refine32[3](_cat([up2(refine32[2](out2)), out[0][0]]))

        preds = [self.head0[idx](out0), self.head1[idx](out1), self.head2[idx](out2),
                 self.head02[idx](_cat([out0, out2])), self.head12[idx](_cat([out1, out2]))]
        if not self.training:
            return tuple(_upsample8(p) for p in preds)
        return tuple(_to_nchw(p) for p in preds)

    def forward_latency(self, size, alpha=True, beta=True, ratio=True):
        """Expected latency of the current architecture distribution from the per-op lookup table
        (model_search.py:361-475): scalar arithmetic on the arch parameters' device."""
        idx = self.arch_idx
        stem = self.stem[idx]
        dev = self._arch("alphas", 0).device
        if alpha:
            alphas = [F.softmax(self._arch("alphas", s), dim=-1) for s in range(3)]
        else:
            alphas = [torch.ones_like(self._arch("alphas", s)).to(dev) * 1. / len(PRIMITIVES) for s in range(3)]
        if beta:
            betas = [None, F.softmax(self._arch("betas", 0), dim=-1), F.softmax(self._arch("betas", 1), dim=-1)]
        else:
            betas = [None, torch.ones_like(self._arch("betas", 0)).to(dev) * 1. / 2, torch.ones_like(self._arch("betas", 1)).to(dev) * 1. / 2]
        if ratio:
            ratios = self.sample_prun_ratio(mode=self.prun_mode if self.prun_mode is not None else self._prun_modes[idx])
        else:
            ratios = self.sample_prun_ratio(mode='max')

        stem_latency = 0
        for op in stem:
            latency, size = op.forward_latency(size)
            stem_latency = stem_latency + latency
        out_prev = [[size, None]]
        latency_total = [[stem_latency, 0], [0, 0], [0, 0]]  # (out, down) per scale

        for i, cells in enumerate(self.cells):
            out, latency = [], []
            for j, cell in enumerate(cells):
                out0 = out1 = down0 = down1 = None
                a = alphas[j][i - j]
                r = self._ratio_triple(i, j, ratios)
                if j == 0:
                    out1, down1 = cell.forward_latency(out_prev[0][0], a, r)
                    out.append((out1[1], down1[1] if down1 is not None else None))
                    latency.append([out1[0], down1[0] if down1 is not None else None])
                elif i == j:
                    out0, down0 = cell.forward_latency(out_prev[j - 1][1], a, r)
                    out.append((out0[1], down0[1] if down0 is not None else None))
                    latency.append([out0[0], down0[0] if down0 is not None else None])
                else:
                    if betas[j][i - j - 1][0] > 0:
                        out0, down0 = cell.forward_latency(out_prev[j - 1][1], a, r)
                    if betas[j][i - j - 1][1] > 0:
                        out1, down1 = cell.forward_latency(out_prev[j][0], a, r)
                    assert (out0 is None and out1 is None) or out0[1] == out1[1]
                    assert (down0 is None and down1 is None) or down0[1] == down1[1]
                    out.append((out0[1], down0[1] if down0 is not None else None))
                    b = betas[j][i - j - 1]
                    latency.append([
                        sum(w * o for w, o in zip(b, [out0[0], out1[0]])),
                        sum(w * d if d is not None else 0 for w, d in zip(b, [down0[0] if down0 is not None else None,
                                                                             down1[0] if down1 is not None else None]))])
            out_prev = out
            for ii, lat in enumerate(latency):
                # layer: i | scale: ii   (kept quirk: the beta row below is indexed with the LAST j of the loop above)
                if ii == 0:
                    if lat[0] is not None: latency_total[ii][0] = latency_total[ii][0] + lat[0]
                    if lat[1] is not None: latency_total[ii][1] = latency_total[ii][0] + lat[1]
                elif i == ii:
                    if lat[0] is not None: latency_total[ii][0] = latency_total[ii - 1][1] + lat[0]
                    if lat[1] is not None: latency_total[ii][1] = latency_total[ii - 1][1] + lat[1]
                else:
                    b = betas[j][i - j - 1]
                    if lat[0] is not None: latency_total[ii][0] = b[1] * latency_total[ii][0] + b[0] * latency_total[ii - 1][1] + lat[0]
                    if lat[1] is not None: latency_total[ii][1] = b[1] * latency_total[ii][0] + b[0] * latency_total[ii - 1][1] + lat[1]
        return sum([latency_total[0][0], latency_total[1][0], latency_total[2][0]])

    def _loss(self, input, target, pretrain=False):
        loss = 0
        if pretrain is not True:
            # "random width": sampled by gumbel softmax
            self.prun_mode = None
            for idx in range(len(self._arch_names)):
                self.arch_idx = idx
                logits = self(input)
                loss = loss + sum(self._criterion(logit, target) for logit in logits)
        if len(self._width_mult_list) > 1:
            modes = ["max", "min"] + (["random", "random"] if pretrain == True else [])
            for mode in modes:
                self.prun_mode = mode
                logits = self(input)
                loss = loss + sum(self._criterion(logit, target) for logit in logits)
        elif pretrain == True and len(self._width_mult_list) == 1:
            self.prun_mode = "max"
            logits = self(input)
            loss = loss + sum(self._criterion(logit, target) for logit in logits)
        return loss

    def _arch_shapes(self, idx):
        num_ops = len(PRIMITIVES)
        num_widths = len(self._width_mult_list) if self._prun_modes[idx] == 'arch_ratio' else 1
        L = self._layers
        return {"alphas": [(L, num_ops), (L - 1, num_ops), (L - 2, num_ops)],
                "betas": [(L - 2, 2), (L - 3, 2)],  # in-degree probs; 0: from down, 1: from keep
                "ratios": [(L - 1, num_widths), (L - 1, num_widths), (L - 2, num_widths)]}

    def _build_arch_parameters(self, idx):
        names = {"alphas": ["alpha_" + str(idx) + "_" + str(s) for s in [0, 1, 2]],
                 "betas": ["beta_" + str(idx) + "_" + str(s) for s in [1, 2]],
                 "ratios": ["ratio_" + str(idx) + "_" + str(s) for s in [0, 1, 2]]}
        shapes = self._arch_shapes(idx)
        for kind in ("alphas", "betas", "ratios"):
            for name, shape in zip(names[kind], shapes[kind]):
                setattr(self, name, nn.Parameter(Variable(1e-3 * torch.ones(*shape), requires_grad=True)))
        params = [getattr(self, n) for kind in ("alphas", "betas", "ratios") for n in names[kind]]
        return names, params

    def _reset_arch_parameters(self, idx):
        shapes = self._arch_shapes(idx)
        for kind in ("alphas", "betas", "ratios"):
            for name, shape in zip(self._arch_names[idx][kind], shapes[kind]):
                getattr(self, name).data = Variable(1e-3 * torch.ones(*shape), requires_grad=True)


# ------------------------------------------------------------------------------------------------------------
def _grad(*ts):
    return torch.is_grad_enabled() and any(t.requires_grad for t in ts)


def _cat(tensors):
    tensors = [F_.to_nhwc_half(t) for t in tensors]
    if _grad(*tensors):
        return AG.cat_channels(tensors)
    N, _, H, W = tensors[0].shape
    out = F_.empty_nhwc(N, sum(t.shape[1] for t in tensors), H, W, tensors[0].device)
    at = 0
    for t in tensors:
        F_.copy_channels(t, out[:, at:at + t.shape[1]])
        at += t.shape[1]
    return out


def _resize2x(t):
    t = F_.to_nhwc_half(t)
    size = (t.shape[2] * 2, t.shape[3] * 2)
    return AG.bilinear(t, size) if _grad(t) else F_.bilinear(t, size)


def _upsample8(p):
    p = F_.to_nhwc_half(p)
    size = (p.shape[2] * 8, p.shape[3] * 8)
    return AG.upsample_logits(p, size) if _grad(p) else F_.upsample_logits(p, size, dtype=torch.float32)


def _to_nchw(p):
    p = F_.to_nhwc_half(p)
    return AG.to_nchw(p) if _grad(p) else F_.to_nchw(p, torch.float32)
