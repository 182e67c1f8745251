"""Training criteria of the reference drivers, restated without host synchronisation so that they can sit between two captured
passes (graphed.py) without stalling the stream.

`ProbOhemCrossEntropy2d` follows tools/seg_opr/loss_opr.py:43-93 (online hard example mining on the softmax probability of the
true class: keep pixels whose probability is <= max(thresh, the min_kept-th smallest probability), then mean cross entropy over
the kept pixels).  The reference sorts all N*H*W probabilities (`argsort`) to read ONE order statistic and branches on
`num_valid` in Python (a GPU->CPU sync per call, 15 calls per search step); here the order statistic is `torch.kthvalue` and the
branches are tensor selects -- same result, pinned against the reference by tests/test_loss_oracle.py / tests/golden/loss.npz.
`distillation_kl` is train/train.py:254-260's KLDivLoss(reduction='mean') term."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class ProbOhemCrossEntropy2d(nn.Module):
    def __init__(self, ignore_label, reduction='mean', thresh=0.6, min_kept=256, down_ratio=1, use_weight=False):
        super(ProbOhemCrossEntropy2d, self).__init__()
        assert reduction == 'mean' and not use_weight, "only the configuration the FasterSeg drivers use is restated"
        self.ignore_label = ignore_label
        self.thresh = float(thresh)
        self.min_kept = int(min_kept)
        self.down_ratio = down_ratio

    def forward(self, pred, target):
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
    logp = F.log_softmax(student_logits.float(), dim=1)
    q = F.softmax(teacher_logits.float(), dim=1)
    return (torch.xlogy(q, q) - q * logp).mean()
