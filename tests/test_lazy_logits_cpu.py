"""Dispatch logic of `losses.LazyLogits` (N1) without a GPU: which of the reference's criterion expressions reach the fused kernels
(train/train.py:254-260: `criterion(logits, target)` with ProbOhemCrossEntropy2d, `nn.KLDivLoss()(F.softmax(s, dim=1).log(),
F.softmax(t, dim=1))`), and that everything else a caller might do with the logits materialises them first and gives what the same
call gives on the F.interpolate'd tensor (train/model_seg.py:357-362).  The fused entry points and the upsample are replaced by
recorders / a torch CPU interpolate here; their arithmetic is checked on the GPU by tests/test_loss_gpu.py."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from fasterseg_b200 import autograd as AG
from fasterseg_b200 import losses, model_seg
from fasterseg_b200.losses import LazyLogits, ProbOhemCrossEntropy2d


@pytest.fixture
def rig(monkeypatch):
    calls = []

    def up(x, size, dtype):
        calls.append(("dense", tuple(size)))
        return F.interpolate(x.float(), size=size, mode="bilinear", align_corners=True).to(dtype)

    def fused_kl(xs, xt, size):
        calls.append(("fused_kl", xs, xt, tuple(size)))
        return torch.tensor(1.25)

    def fused_ohem(x, target, size, ignore_label, thresh, min_kept):
        calls.append(("fused_ohem", x, tuple(size), ignore_label, thresh, min_kept))
        return torch.tensor(2.5)

    monkeypatch.setattr(model_seg, "_upsample_logits", up)
    monkeypatch.setattr(AG, "fused_kl", fused_kl)
    monkeypatch.setattr(AG, "fused_ohem_ce", fused_ohem)
    g = torch.Generator().manual_seed(0)
    s = torch.randn(2, 19, 4, 8, generator=g)
    t = torch.randn(2, 19, 4, 8, generator=g)
    return calls, LazyLogits(s, (32, 64)), LazyLogits(t, (32, 64)), s, t


def _dense(x):
    return F.interpolate(x, size=(32, 64), mode="bilinear", align_corners=True)


def test_shape_protocol(rig):
    _, ls, _, _, _ = rig
    assert ls.shape == torch.Size((2, 19, 32, 64)) and ls.size_() == ls.shape and ls.size_(1) == 19 and ls.float() is ls
    d = ls.detach()
    assert isinstance(d, LazyLogits) and d.size == (32, 64) and not d.lowres.requires_grad


def test_reference_kl_expression_reaches_the_fused_kernel(rig):
    calls, ls, lt, s, t = rig
    loss = nn.KLDivLoss()(F.softmax(ls, dim=1).log(), F.softmax(lt, dim=1))            # train/train.py:259, verbatim
    assert float(loss) == 1.25 and [c[0] for c in calls] == ["fused_kl"]
    assert calls[0][1] is s and calls[0][3] == (32, 64) and torch.equal(calls[0][2], t) and not calls[0][2].requires_grad
    calls.clear()
    assert float(F.kl_div(F.log_softmax(ls, dim=1), torch.softmax(lt, 1), reduction="mean")) == 1.25      # the same thing, spelled differently
    assert float(F.kl_div(F.log_softmax(ls, dim=-3), F.softmax(lt, dim=1))) == 1.25
    assert [c[0] for c in calls] == ["fused_kl", "fused_kl"]
    assert float(losses.distillation_kl(ls, lt)) == 1.25


@pytest.mark.parametrize("variant", ["batchmean", "sum", "log_target", "other_dim", "two_logs", "legacy_size_average"])
def test_other_kl_spellings_materialise_and_match_the_dense_result(rig, variant):
    calls, ls, lt, s, t = rig
    ds, dt = _dense(s), _dense(t)
    if variant in ("batchmean", "sum"):
        got = F.kl_div(F.softmax(ls, dim=1).log(), F.softmax(lt, dim=1), reduction=variant)
        want = F.kl_div(F.softmax(ds, dim=1).log(), F.softmax(dt, dim=1), reduction=variant)
    elif variant == "log_target":
        got = F.kl_div(F.log_softmax(ls, dim=1), F.log_softmax(lt, dim=1), log_target=True)
        want = F.kl_div(F.log_softmax(ds, dim=1), F.log_softmax(dt, dim=1), log_target=True)
    elif variant == "other_dim":
        got = F.kl_div(F.softmax(ls, dim=2).log(), F.softmax(lt, dim=2))
        want = F.kl_div(F.softmax(ds, dim=2).log(), F.softmax(dt, dim=2))
    elif variant == "two_logs":
        got = F.kl_div(F.log_softmax(ls, dim=1), F.log_softmax(lt, dim=1))             # target given as log-probabilities without the flag
        want = F.kl_div(F.log_softmax(ds, dim=1), F.log_softmax(dt, dim=1))
    else:
        with pytest.warns(UserWarning):
            got = F.kl_div(F.log_softmax(ls, dim=1), F.softmax(lt, dim=1), size_average=False)
        with pytest.warns(UserWarning):
            want = F.kl_div(F.log_softmax(ds, dim=1), F.softmax(dt, dim=1), size_average=False)
    assert "fused_kl" not in [c[0] for c in calls] and "dense" in [c[0] for c in calls]
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-7, equal_nan=True)      # "two_logs" is a misuse: nan on both sides


def test_ohem_criterion_takes_the_fused_path_on_lazy_logits_only(rig):
    calls, ls, _, s, _ = rig
    tgt = torch.randint(0, 19, (2, 32, 64), generator=torch.Generator().manual_seed(3))
    tgt[0, :3] = 255
    crit = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=100)
    assert float(crit(ls, tgt)) == 2.5
    assert calls[-1][0] == "fused_ohem" and calls[-1][1] is s and calls[-1][2:] == ((32, 64), 255, 0.7, 100)
    calls.clear()
    dense_loss = crit(_dense(s), tgt)                       # a materialised tensor: the sync-free restatement, pinned by test_loss_oracle.py
    assert not calls and dense_loss.dim() == 0 and float(dense_loss) > 0


def test_anything_else_sees_the_materialised_logits(rig):
    calls, ls, lt, s, t = rig
    ds = _dense(s)
    assert torch.equal(torch.argmax(ls, dim=1), ds.argmax(1))
    tgt = torch.randint(0, 19, (2, 32, 64), generator=torch.Generator().manual_seed(4))
    assert torch.allclose(F.cross_entropy(ls, tgt), F.cross_entropy(ds, tgt))
    assert torch.allclose(torch.add(ls, 1.0), ds + 1.0)
    sm = F.softmax(ls, dim=1)
    assert torch.allclose(sm.dense(), F.softmax(ds, dim=1)) and torch.allclose(sm.log().dense(), F.log_softmax(ds, dim=1), atol=1e-6)
    assert torch.allclose(torch.sum(sm, dim=1), torch.ones(2, 32, 64), atol=1e-5)
    assert all(c[0] == "dense" for c in calls)
    lt_small = LazyLogits(t, (16, 32))
    with pytest.raises(AssertionError):
        losses.distillation_kl(ls, lt_small)                 # student and teacher must be asked for at the same label resolution
