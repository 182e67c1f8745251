"""Host logic of the fused criteria (N1; fasterseg_b200/autograd.py FusedOhemCEFn / FusedKLFn, losses.py) on the CPU stand-in kernels:
the decisions the autograd functions make around the kernels -- threshold = max(k-th smallest probability, thresh) in log space, no
mining when fewer valid pixels than min_kept, min_kept = 0, nothing valid at all, the 1 / count and 1 / numel coefficients, GRAD_SCALE
-- against the oracle pinned to the unmodified reference (tools/seg_opr/loss_opr.py:63-93, train/train.py:254-260; goldens
tests/golden/loss.npz) and against the materialised x8 / x16 path.  The kernels themselves: tests/test_loss_gpu.py."""
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from fasterseg_b200 import autograd as AG
from fasterseg_b200 import functional as F_
from fasterseg_b200.losses import LazyLogits, ProbOhemCrossEntropy2d, distillation_kl
from oracle import fasterseg_oracle as orc
from oracle import make_golden_loss as mk
from tests import cpu_backend


@pytest.fixture(autouse=True)
def _cpu_backend():
    with cpu_backend.installed():
        yield


def _lowres(x_nchw_f32):
    N, C, H, W = x_nchw_f32.shape
    buf = F_.empty_nhwc(N, C, H, W, "cpu")
    buf.copy_(x_nchw_f32.half())
    return buf.requires_grad_(True)


def _grad_close(got, ref, what):
    scale = float(ref.abs().max()) + 1e-30
    err = float((got.double() - ref.double()).abs().max())
    assert err <= 2e-3 * scale, "%s: max |diff| %.3e vs max |ref| %.3e" % (what, err, scale)


@pytest.mark.parametrize("name", sorted(mk.OHEM_CASES))
def test_fused_ohem_function_follows_every_branch_of_the_reference(name):
    pred, tgt, thresh, min_kept = mk.ohem_inputs(name)
    pred = pred.half().float()
    ref_in = pred.clone().requires_grad_(True)
    want = orc.ohem_cross_entropy(ref_in, tgt, ignore_label=255, thresh=thresh, min_kept=min_kept)
    x = _lowres(pred)
    loss = ProbOhemCrossEntropy2d(ignore_label=255, thresh=thresh, min_kept=min_kept)(LazyLogits(x, pred.shape[2:]), tgt)
    assert "FusedOhemCE" in type(loss.grad_fn).__name__
    if math.isnan(float(want)):
        assert math.isnan(float(loss))
        return
    assert float(loss) == pytest.approx(float(want), rel=2e-6)
    want.backward()
    loss.backward()
    _grad_close(x.grad.float() / AG.GRAD_SCALE, ref_in.grad, name)


@pytest.mark.parametrize("factor,hw,batch", [(8, (8, 16), 2), (16, (4, 6), 2)])
def test_fused_ohem_function_equals_the_materialised_criterion(factor, hw, batch):
    g = torch.Generator().manual_seed(factor)
    h, w = hw
    H, W = h * factor, w * factor
    low = (torch.randn(batch, 19, h, w, generator=g) * 2.5).half().float()
    tgt = torch.randint(0, 19, (batch, H, W), generator=g)
    tgt[torch.rand(tgt.shape, generator=g) < 0.06] = 255
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=batch * H * W // 16)
    dense_in = low.clone().requires_grad_(True)
    want = crit(F.interpolate(dense_in, size=(H, W), mode="bilinear", align_corners=True), tgt)
    want.backward()
    x = _lowres(low)
    loss = 0.2 * crit(LazyLogits(x, (H, W)), tgt)             # the drivers weight the x16 / x32 terms (train/train.py:256-257)
    loss.backward()
    assert float(loss) == pytest.approx(0.2 * float(want), rel=5e-6)
    _grad_close(x.grad.float() / AG.GRAD_SCALE, 0.2 * dense_in.grad, "x%d" % factor)


@pytest.mark.parametrize("name", sorted(mk.KL_CASES))
def test_fused_kl_function_matches_the_pinned_oracle(name):
    student, teacher = mk.kl_inputs(name)
    student, teacher = student.half().float(), teacher.half().float()
    ref_in = student.clone().requires_grad_(True)
    want = orc.distill_kl(ref_in, teacher)
    want.backward()
    xs, xt = _lowres(student), _lowres(teacher).detach()
    loss = nn.KLDivLoss()(F.softmax(LazyLogits(xs, student.shape[2:]), dim=1).log(), F.softmax(LazyLogits(xt, student.shape[2:]), dim=1))
    assert "FusedKL" in type(loss.grad_fn).__name__
    loss.backward()
    assert float(loss) == pytest.approx(float(want), rel=1e-5)
    _grad_close(xs.grad.float() / AG.GRAD_SCALE, ref_in.grad, name)
    assert float(distillation_kl(LazyLogits(xs, student.shape[2:]), LazyLogits(xt, student.shape[2:]))) == pytest.approx(float(want), rel=1e-5)


def test_all_terms_of_the_distillation_loss_accumulate_into_one_low_resolution_gradient():
    """train/train.py:254-260: three OHEM terms and the KL term all hang off the student's logits8; their gradients must add up"""
    g = torch.Generator().manual_seed(9)
    low = (torch.randn(2, 19, 6, 10, generator=g) * 2).half().float()
    low_t = (torch.randn(2, 19, 6, 10, generator=g) * 2).half().float()
    size = (48, 80)
    tgt = torch.randint(0, 19, (2,) + size, generator=g)
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=500)
    dense_in = low.clone().requires_grad_(True)
    ds = F.interpolate(dense_in, size=size, mode="bilinear", align_corners=True)
    dt = F.interpolate(low_t, size=size, mode="bilinear", align_corners=True)
    want = crit(ds, tgt) + nn.KLDivLoss()(F.softmax(ds, dim=1).log(), F.softmax(dt, dim=1))
    want.backward()
    x, xt = _lowres(low), _lowres(low_t).detach()
    lazy = LazyLogits(x, size)
    loss = crit(lazy, tgt) + nn.KLDivLoss()(F.softmax(lazy, dim=1).log(), F.softmax(LazyLogits(xt, size), dim=1))
    loss.backward()
    assert float(loss) == pytest.approx(float(want), rel=1e-5)
    _grad_close(x.grad.float() / AG.GRAD_SCALE, dense_in.grad, "ohem + kl")


def test_distillation_step_with_lazy_logits_matches_the_materialised_path():
    """the whole wiring of train/train.py:243-269 on the CPU stand-in: teacher in eval, student in train mode, `lazy_logits` on both ->
    the heads hand out LazyLogits, the three OHEM terms and the KL term take the fused path, gradients reach every student parameter;
    same loss and gradients as with materialised logits (two roundings of the same mathematics, amplified by the BatchNorm chains --
    DESIGN section 4 -- hence the loose gate on the median)"""
    from bench import synth_weights_
    from tests.test_boundary_cpu import _build_student
    g = torch.Generator().manual_seed(0)
    B, Hh, Ww = 2, 64, 128
    x = torch.randn(B, 3, Hh, Ww, generator=g)
    t = torch.randint(0, 19, (B, Hh, Ww), generator=g)
    t[torch.rand(t.shape, generator=g) < 0.05] = 255
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=B * Hh * Ww // 16)
    results = []
    for lazy in (False, True):
        teacher, _ = _build_student(0)
        synth_weights_(teacher, 1)
        student, _ = _build_student(1, training=True)
        synth_weights_(student, 2)
        teacher.lazy_logits = lazy
        student.lazy_logits = lazy
        with torch.no_grad():
            tl = teacher.eval()(x)
        l8, l16, l32 = student(x)
        assert isinstance(l8, LazyLogits) == lazy and isinstance(tl, LazyLogits) == lazy
        loss = crit(l8, t) + 0.2 * crit(l16, t) + 0.2 * crit(l32, t) + distillation_kl(l8, tl)
        loss.backward()
        grads = {k: p.grad.detach().float().clone() for k, p in student.named_parameters() if p.grad is not None}
        missing = [k for k, p in student.named_parameters() if p.grad is None]
        results.append((float(loss.detach()), grads, missing))
    (l0, g0, m0), (l1, g1, m1) = results
    assert l1 == pytest.approx(l0, rel=2e-3), (l0, l1)
    assert set(g0) == set(g1) and m0 == m1 and len(g0) > 100
    rel = sorted(float((g1[k] - g0[k]).norm() / (g0[k].norm() + 1e-12)) for k in g0)
    assert rel[len(rel) // 2] < 2e-2, "median relative gradient difference %.3e" % rel[len(rel) // 2]
