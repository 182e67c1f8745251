"""GPU parity of the fused criteria (N1, csrc/loss.cu): OHEM cross entropy and KL distillation evaluated from LOW-RESOLUTION logits.

Oracles: (a) oracle/fasterseg_oracle.py's restatement of tools/seg_opr/loss_opr.py:63-93 / train/train.py:254-260, itself pinned to
the unmodified reference by tests/golden/loss.npz; (b) the dense path (F.interpolate + the same criterion on the materialised
label-resolution logits) for the x8 / x16 / x32 cases.  Inputs are rounded to fp16 once (the kernels read fp16 NHWC logits); the
loss is compared at fp32 accuracy, the gradient -- which leaves the kernel as fp16 x GRAD_SCALE -- at fp16 accuracy."""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import fasterseg_oracle as orc
from oracle import make_golden_loss as mk

pytestmark = pytest.mark.gpu


def _nhwc(x):
    return x.cuda().half().contiguous(memory_format=torch.channels_last)


def _lowres(x_nchw_f32):
    """CPU NCHW fp32 -> our NHWC fp16 logits buffer (channel stride rounded up to 8) as an autograd leaf"""
    from fasterseg_b200 import functional as F_
    N, C, H, W = x_nchw_f32.shape
    buf = F_.empty_nhwc(N, C, H, W, "cuda")
    buf.copy_(x_nchw_f32.cuda().half())
    return buf.requires_grad_(True)


def _grad_close(got, ref, what):
    ref = ref.double()
    got = got.double()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item()
    assert err <= 2e-3 * scale, "%s: max |diff| %.3e vs max |ref| %.3e" % (what, err, scale)


@pytest.mark.parametrize("n", [1, 7, 1000, 4096 * 37 + 11, 1 << 21])
def test_kth_smallest_is_exact(n):
    from fasterseg_b200 import functional as F_
    g = torch.Generator().manual_seed(n)
    x = -torch.rand(n, generator=g).log()            # positive, heavy tail
    x = torch.where(torch.rand(n, generator=g) < 0.5, -x, x)
    if n > 100:
        x[::17] = 0.0                                # ties (the ignored pixels of the loss carry exactly 0)
        x[5::97] = x[3]                              # more ties
    xs = torch.sort(x).values
    xc = x.cuda().contiguous()
    for k in sorted(k for k in {1, 2, n // 3 + 1, n // 2 + 1, n - 1, n} if 1 <= k <= n):
        got = float(F_.kth_smallest(xc, k))
        assert got == float(xs[k - 1]), (n, k, got, float(xs[k - 1]))


@pytest.mark.parametrize("name", sorted(mk.OHEM_CASES))
def test_fused_ohem_matches_the_pinned_oracle_at_identity_scale(name):
    """label resolution == logits resolution: the upsample is the identity, the criterion must equal the reference's on the same
    (fp16-rounded) logits -- every decision branch of loss_opr.py:63-93"""
    from fasterseg_b200 import autograd as AG
    from fasterseg_b200.losses import LazyLogits, ProbOhemCrossEntropy2d
    pred, tgt, thresh, min_kept = mk.ohem_inputs(name)
    pred = pred.half().float()
    ref_in = pred.clone().requires_grad_(True)
    want = orc.ohem_cross_entropy(ref_in, tgt, ignore_label=255, thresh=thresh, min_kept=min_kept)
    x = _lowres(pred)
    loss = ProbOhemCrossEntropy2d(ignore_label=255, thresh=thresh, min_kept=min_kept)(LazyLogits(x, pred.shape[2:]), tgt.cuda())
    if math.isnan(float(want)):
        assert math.isnan(float(loss))
        return
    assert float(loss) == pytest.approx(float(want), rel=2e-6)
    want.backward()
    loss.backward()
    _grad_close(x.grad.float().cpu() / AG.GRAD_SCALE, ref_in.grad, name)


@pytest.mark.parametrize("factor,hw,batch", [(8, (16, 32), 3), (16, (8, 12), 2), (32, (5, 7), 2)])
def test_fused_ohem_matches_the_dense_path(factor, hw, batch):
    """x8 / x16 / x32 heads (train/model_seg.py:357-362): fused criterion vs F.interpolate + criterion on the materialised logits"""
    from fasterseg_b200 import autograd as AG
    from fasterseg_b200.losses import LazyLogits, ProbOhemCrossEntropy2d
    g = torch.Generator().manual_seed(factor)
    h, w = hw
    H, W = h * factor, w * factor
    low = (torch.randn(batch, 19, h, w, generator=g) * 2.5).half().float()
    tgt = torch.randint(0, 19, (batch, H, W), generator=g)
    tgt[torch.rand(tgt.shape, generator=g) < 0.06] = 255
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=batch * H * W // 16)
    dense_in = low.clone().cuda().requires_grad_(True)
    dense = F.interpolate(dense_in, size=(H, W), mode="bilinear", align_corners=True)
    want = crit(dense, tgt.cuda())
    want.backward()
    x = _lowres(low)
    loss = crit(LazyLogits(x, (H, W)), tgt.cuda())
    loss.backward()
    assert float(loss) == pytest.approx(float(want), rel=5e-6)
    _grad_close(x.grad.float().cpu() / AG.GRAD_SCALE, dense_in.grad.cpu(), "x%d" % factor)


@pytest.mark.parametrize("name", sorted(mk.KL_CASES))
def test_fused_kl_matches_the_pinned_oracle_at_identity_scale(name):
    from fasterseg_b200 import autograd as AG
    from fasterseg_b200.losses import LazyLogits, distillation_kl
    student, teacher = mk.kl_inputs(name)
    student, teacher = student.half().float(), teacher.half().float()
    ref_in = student.clone().requires_grad_(True)
    want = orc.distill_kl(ref_in, teacher)
    want.backward()
    xs, xt = _lowres(student), _lowres(teacher).detach()
    loss = distillation_kl(LazyLogits(xs, student.shape[2:]), LazyLogits(xt, student.shape[2:]))
    loss.backward()
    assert float(loss) == pytest.approx(float(want), rel=1e-5)
    _grad_close(xs.grad.float().cpu() / AG.GRAD_SCALE, ref_in.grad, name)


def test_reference_kl_expression_takes_the_fused_path_and_matches_dense():
    """train/train.py:254-260 verbatim: nn.KLDivLoss()(F.softmax(student, dim=1).log(), F.softmax(teacher, dim=1)) on LazyLogits"""
    from fasterseg_b200 import autograd as AG
    from fasterseg_b200.losses import LazyLogits
    g = torch.Generator().manual_seed(5)
    low_s = (torch.randn(2, 19, 12, 20, generator=g) * 2).half().float()
    low_t = (torch.randn(2, 19, 12, 20, generator=g) * 2).half().float()
    size = (96, 160)
    dense_in = low_s.clone().cuda().requires_grad_(True)
    ds = F.interpolate(dense_in, size=size, mode="bilinear", align_corners=True)
    dt = F.interpolate(low_t.cuda(), size=size, mode="bilinear", align_corners=True)
    want = nn.KLDivLoss()(F.softmax(ds, dim=1).log(), F.softmax(dt, dim=1))
    want.backward()
    xs, xt = _lowres(low_s), _lowres(low_t).detach()
    loss = nn.KLDivLoss()(F.softmax(LazyLogits(xs, size), dim=1).log(), F.softmax(LazyLogits(xt, size), dim=1))
    assert loss.grad_fn is not None and "FusedKL" in type(loss.grad_fn).__name__
    loss.backward()
    assert float(loss) == pytest.approx(float(want), rel=2e-5)
    _grad_close(xs.grad.float().cpu() / AG.GRAD_SCALE, dense_in.grad.cpu(), "kl x8")


def test_student_train_step_with_lazy_logits_matches_the_materialised_path():
    """distillation-style criterion (3 OHEM terms + KL) on the student in train mode: lazy (fused) vs materialised logits"""
    from bench import synth_weights_
    from fasterseg_b200 import zoo
    from fasterseg_b200.losses import ProbOhemCrossEntropy2d, distillation_kl
    torch.manual_seed(0)
    B, H, W = 2, 128, 256
    x = torch.randn(B, 3, H, W, device="cuda")
    t = torch.randint(0, 19, (B, H, W), device="cuda")
    t[torch.rand(t.shape, device="cuda") < 0.05] = 255
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=B * H * W // 16)
    results = []
    for lazy in (False, True):
        teacher = zoo.build_network(0).cuda().eval()
        synth_weights_(teacher, 1)
        student = zoo.build_network(1, training=True).cuda().train()
        synth_weights_(student, 2)
        teacher.lazy_logits = lazy
        student.lazy_logits = lazy
        with torch.no_grad():
            tl = teacher(x)
        l8, l16, l32 = student(x)
        loss = crit(l8, t) + 0.2 * crit(l16, t) + 0.2 * crit(l32, t) + distillation_kl(l8, tl)
        loss.backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().float().clone() for k, p in student.named_parameters() if p.grad is not None}
        results.append((float(loss), grads))
    (l0, g0), (l1, g1) = results
    assert l1 == pytest.approx(l0, rel=2e-3), (l0, l1)
    assert set(g0) == set(g1)
    rel = sorted(float((g1[k] - g0[k]).norm() / (g0[k].norm() + 1e-12)) for k in g0)
    assert rel[len(rel) // 2] < 2e-2, "median relative gradient difference %.3e" % rel[len(rel) // 2]
