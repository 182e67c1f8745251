#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- goldens for the callers' loss path ("next" row N1 of SURVEY section 8f): the reference's
`ProbOhemCrossEntropy2d` (tools/seg_opr/loss_opr.py:43-93, used as train/train.py:79-81 builds it: thresh 0.7, min_kept =
batch * H * W // 16) and the distillation term of train/train.py:254-260 (`KLDivLoss()(log softmax(student), softmax(teacher))`),
evaluated by the UNMODIFIED reference on seeded inputs; loss values and the gradient w.r.t. the logits are stored
(strided) in tests/golden/loss.npz.   Run in the build container:  python oracle/make_golden_loss.py"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness  # noqa: E402

# name -> (batch, H, W, seed, thresh, min_kept, fraction of ignored labels, logit scale)
OHEM_CASES = {
    "train_like": (2, 64, 128, 11, 0.7, 2 * 64 * 128 // 16, 0.05, 2.0),       # some pixels above the threshold are dropped
    "kth_above_thresh": (2, 32, 64, 12, 0.3, 2 * 32 * 64 // 4, 0.05, 4.0),     # the min_kept-th smallest prob > thresh
    "few_valid": (1, 16, 32, 13, 0.7, 1000, 0.9, 2.0),                        # min_kept > num_valid: plain CE
    "all_ignored": (1, 8, 16, 14, 0.7, 16, 1.0, 2.0),                         # no valid pixel at all
    "no_min_kept": (2, 16, 32, 15, 0.6, 0, 0.1, 3.0),                         # min_kept = 0: fixed threshold
}
KL_CASES = {"kl_small": (2, 32, 64, 21), "kl_train_like": (3, 64, 128, 22)}


def ohem_inputs(name):
    b, h, w, seed, thresh, min_kept, p_ignore, scale = OHEM_CASES[name]
    rs = np.random.RandomState(seed)
    pred = torch.from_numpy((rs.standard_normal((b, 19, h, w)) * scale).astype(np.float32))
    tgt = rs.randint(0, 19, size=(b, h, w)).astype(np.int64)
    tgt[rs.uniform(size=tgt.shape) < p_ignore] = 255
    return pred, torch.from_numpy(tgt), thresh, min_kept


def kl_inputs(name):
    b, h, w, seed = KL_CASES[name]
    rs = np.random.RandomState(seed)
    mk = lambda: torch.from_numpy((rs.standard_normal((b, 19, h, w)) * 2.0).astype(np.float32))
    return mk(), mk()


def main():
    ns = ref_harness.load_reference("train", "seg_opr.loss_opr")
    Ohem = ns.modules["seg_opr.loss_opr"].ProbOhemCrossEntropy2d
    rec = {}
    for name in OHEM_CASES:
        pred, tgt, thresh, min_kept = ohem_inputs(name)
        pred.requires_grad_(True)
        crit = Ohem(ignore_label=255, thresh=thresh, min_kept=min_kept, use_weight=False)
        loss = crit(pred, tgt.clone())
        rec[name + "/loss"] = np.array([float(loss.detach())], dtype=np.float64)
        if torch.isfinite(loss):
            loss.backward()
            rec[name + "/grad.s4"] = pred.grad.numpy()[:, :, ::4, ::4].copy()
        print(name, float(loss))
    for name in KL_CASES:
        student, teacher = kl_inputs(name)
        student.requires_grad_(True)
        loss = nn.KLDivLoss()(F.softmax(student, dim=1).log(), F.softmax(teacher, dim=1))   # train/train.py:79,260
        loss.backward()
        rec[name + "/loss"] = np.array([float(loss.detach())], dtype=np.float64)
        rec[name + "/grad.s4"] = student.grad.numpy()[:, :, ::4, ::4].copy()
        print(name, float(loss))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "loss.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
