"""Training criteria of the reference drivers, restated without host synchronisation so that they can sit between two captured
passes (graphed.py) without stalling the stream.

`ProbOhemCrossEntropy2d` follows tools/seg_opr/loss_opr.py:43-93 (online hard example mining on the softmax probability of the
true class: keep pixels whose probability is <= max(thresh, the min_kept-th smallest probability), then mean cross entropy over
the kept pixels).  The reference sorts all N*H*W probabilities (`argsort`) to read ONE order statistic and branches on
`num_valid` in Python (a GPU->CPU sync per call, 15 calls per search step); here the order statistic is `torch.kthvalue` and the
branches are tensor selects -- same result, pinned against the reference by tests/test_loss_oracle.py / tests/golden/loss.npz.
`distillation_kl` is train/train.py:254-260's KLDivLoss(reduction='mean') term.

N1, fused path: with `model.lazy_logits = True` the student / teacher return `LazyLogits` -- the LOW-RESOLUTION head
output plus the size it would be upsampled to -- instead of materialised label-resolution tensors.  The criteria below (and, through
`__torch_function__`, the reference's own `nn.KLDivLoss()(F.softmax(s, dim=1).log(), F.softmax(t, dim=1))` expression) recognise
them and run csrc/loss.cu: upsample -> log-softmax -> OHEM / KL in one pass per direction, no 478 MB logits, no argsort, no host
synchronisation.  Any other use of a LazyLogits materialises it (`.dense()`), so unmodified callers keep working."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class LazyLogits:
    """Upsampled logits that have not been materialised: `lowres` is the NHWC fp16 head output (autograd-tracked), `size` the label
    resolution of train/model_seg.py:357-362's F.interpolate(..., mode='bilinear', align_corners=True)."""

    def __init__(self, lowres, size, dtype=torch.float32):
        self.lowres, self.size, self.dtype = lowres, (int(size[0]), int(size[1])), dtype

    @property
    def shape(self):
        return torch.Size((self.lowres.shape[0], self.lowres.shape[1]) + self.size)

    def size_(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dense(self):
        from .model_seg import _upsample_logits
        return _upsample_logits(self.lowres, self.size, self.dtype)

    def float(self):
        return self

    def detach(self):
        return LazyLogits(self.lowres.detach(), self.size, self.dtype)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (F.softmax, torch.softmax) and isinstance(args[0], LazyLogits) and _class_dim(args, kwargs):
            return _LazySoftmax(args[0], log=False)
        if func in (F.log_softmax, torch.log_softmax) and isinstance(args[0], LazyLogits) and _class_dim(args, kwargs):
            return _LazySoftmax(args[0], log=True)
        return func(*_densify(args), **{k: _densify((v,))[0] for k, v in kwargs.items()})


class _LazySoftmax:
    """softmax / log_softmax over the class axis of a LazyLogits, still not materialised"""

    def __init__(self, logits, log):
        self.logits, self.is_log = logits, log

    def log(self):
        return _LazySoftmax(self.logits, True) if not self.is_log else self.dense().log()

    def dense(self):
        d = self.logits.dense().float()
        return F.log_softmax(d, dim=1) if self.is_log else F.softmax(d, dim=1)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is F.kl_div and len(args) >= 2 and isinstance(args[0], _LazySoftmax) and isinstance(args[1], _LazySoftmax) \
                and args[0].is_log and not args[1].is_log and not kwargs.get("log_target", False) \
                and _kl_reduction(args, kwargs) == "mean" and args[0].logits.size == args[1].logits.size:
            return distillation_kl(args[0].logits, args[1].logits)
        return func(*_densify(args), **{k: _densify((v,))[0] for k, v in kwargs.items()})


def _class_dim(args, kwargs):
    dim = kwargs.get("dim", args[1] if len(args) > 1 else None)
    return dim in (1, -3)


def _kl_reduction(args, kwargs):
    # F.kl_div(input, target, size_average=None, reduce=None, reduction='mean', log_target=False); nn.KLDivLoss() passes reduction='mean'
    if kwargs.get("size_average") is not None or kwargs.get("reduce") is not None:
        return None
    return kwargs.get("reduction", args[4] if len(args) > 4 else "mean")


def _densify(seq):
    return tuple(a.dense() if isinstance(a, (LazyLogits, _LazySoftmax)) else a for a in seq)


class ProbOhemCrossEntropy2d(nn.Module):
    def __init__(self, ignore_label, reduction='mean', thresh=0.6, min_kept=256, down_ratio=1, use_weight=False):
        super(ProbOhemCrossEntropy2d, self).__init__()
        assert reduction == 'mean' and not use_weight, "only the configuration the FasterSeg drivers use is restated"
        self.ignore_label = ignore_label
        self.thresh = float(thresh)
        self.min_kept = int(min_kept)
        self.down_ratio = down_ratio

    def forward(self, pred, target):
        if isinstance(pred, LazyLogits):      # fused path (csrc/loss.cu): criterion straight from the low-resolution logits
            from . import autograd as AG
            return AG.fused_ohem_ce(pred.lowres, target, pred.size, self.ignore_label, self.thresh, self.min_kept)
        b, c, h, w = pred.shape
        target = target.reshape(-1)
        valid = target.ne(self.ignore_label)
        tgt = target * valid.long()
        num_valid = valid.sum()
        logp = F.log_softmax(pred.float(), dim=1)
        logp_t = logp.permute(0, 2, 3, 1).reshape(-1, c).gather(1, tgt.view(-1, 1)).view(-1)     # log prob of the true class
        if self.min_kept > 0:
            with torch.no_grad():
                prob_t = logp_t.exp().masked_fill(~valid, 1.0)
                k = min(prob_t.numel(), self.min_kept)
                kth = torch.kthvalue(prob_t, k).values
                threshold = torch.clamp(kth, min=self.thresh)            # kth if kth > thresh else thresh
                mining = (num_valid >= self.min_kept) & (num_valid > 0)  # reference: no mining when fewer valid than min_kept
                valid = valid & (prob_t.le(threshold) | ~mining)
        kept = valid.to(logp_t.dtype)
        return -(logp_t * kept).sum() / kept.sum()


def distillation_kl(student_logits, teacher_logits):
    """nn.KLDivLoss(reduction='mean')(log_softmax(student, 1), softmax(teacher, 1)) -- the element-wise mean, like the reference"""
    if isinstance(student_logits, LazyLogits) and isinstance(teacher_logits, LazyLogits):
        from . import autograd as AG
        assert student_logits.size == teacher_logits.size
        return AG.fused_kl(student_logits.lowres, teacher_logits.lowres.detach(), student_logits.size)
    student_logits, teacher_logits = _densify((student_logits, teacher_logits))
    logp = F.log_softmax(student_logits.float(), dim=1)
    q = F.softmax(teacher_logits.float(), dim=1)
    return (torch.xlogy(q, q) - q * logp).mean()
