"""Host logic of the flat step tail (fasterseg_b200/optim.py) that needs no GPU: whenever the gradients do NOT come out of a captured
`_loss` the replacements the launcher installs for torch.optim.SGD / nn.utils.clip_grad_norm_ (search/train_search.py:94-101,249-250;
train/train.py:150,269) must behave exactly like torch's own -- same parameters, same return value -- and the whole-model shortcut must
recognise `model.parameters()` without consuming it.  The kernels themselves are checked by tests/test_optim_gpu.py."""
import weakref

import numpy as np
import pytest
import torch
import torch.nn as nn

from fasterseg_b200 import graphed
from fasterseg_b200 import optim as FO


def _mlp(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))


@pytest.mark.parametrize("kw", [dict(momentum=0.9, weight_decay=5e-4), dict(momentum=0.0), dict(momentum=0.9, nesterov=True),
                                dict(momentum=0.5, dampening=0.1)])
def test_flat_sgd_without_a_flat_buffer_is_torch_sgd(kw):
    a, b = _mlp(3), _mlp(3)
    oa, ob = FO.FlatSGD(a.parameters(), lr=0.1, **kw), torch.optim.SGD(b.parameters(), lr=0.1, **kw)
    x = torch.randn(7, 6)
    for _ in range(4):
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            m(x).pow(2).mean().backward()
            o.step()
    assert oa.flat_steps == 0
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p, q)


def test_lr_scheduler_drives_the_wrapped_optimizer():
    a, b = _mlp(4), _mlp(4)
    oa, ob = FO.FlatSGD(a.parameters(), lr=0.2, momentum=0.9), torch.optim.SGD(b.parameters(), lr=0.2, momentum=0.9)
    sa, sb = (torch.optim.lr_scheduler.ExponentialLR(o, 0.5) for o in (oa, ob))      # train_search.py:101 uses ExponentialLR
    x = torch.randn(4, 6)
    for _ in range(3):
        for m, o, s in ((a, oa, sa), (b, ob, sb)):
            o.zero_grad()
            m(x).sum().backward()
            o.step()
            s.step()
    assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"] == pytest.approx(0.2 * 0.5 ** 3)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p, q)


@pytest.mark.parametrize("as_generator", [True, False])
@pytest.mark.parametrize("norm_type", [2.0, 1.0, float("inf")])
def test_clip_without_a_flat_buffer_is_torch_clip(as_generator, norm_type):
    a, b = _mlp(5), _mlp(5)
    x = torch.randn(9, 6) * 4
    for m in (a, b):
        m(x).pow(2).sum().backward()
    got = FO.clip_grad_norm_(a.parameters() if as_generator else list(a.parameters()), 0.3, norm_type=norm_type)
    want = torch.nn.utils.clip_grad_norm_(b.parameters(), 0.3, norm_type=norm_type)
    assert torch.equal(got, want)
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.equal(p.grad, q.grad)


def test_single_tensor_argument():
    w = nn.Parameter(torch.ones(4))
    w.grad = torch.full((4,), 2.0)
    n = FO.clip_grad_norm_(w, 1.0)
    assert float(n) == pytest.approx(4.0)
    assert torch.allclose(w.grad, torch.full((4,), 0.5), rtol=1e-5)


def test_install_swaps_and_uninstall_restores():
    sgd, clip = torch.optim.SGD, torch.nn.utils.clip_grad_norm_
    FO.install()
    try:
        assert torch.optim.SGD is FO.FlatSGD and torch.nn.utils.clip_grad_norm_ is FO.clip_grad_norm_
        m = _mlp(6)
        o = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)      # the reference's constructor call shape
        assert isinstance(o, FO.FlatSGD) and isinstance(o, torch.optim.Optimizer)
    finally:
        FO.uninstall()
    assert torch.optim.SGD is sgd and torch.nn.utils.clip_grad_norm_ is clip


class _FakeFlat:
    def __init__(self, model):
        self.model_ref = weakref.ref(model)


def test_whole_model_generator_is_recognised_without_being_consumed():
    m = _mlp(7)
    assert FO._whole_model_generator(list(m.parameters())) is None            # not a generator
    assert FO._whole_model_generator(p for p in m.parameters()) is None       # some other generator
    assert FO._whole_model_generator(m.parameters()) is None                  # no flat buffer registered for this model
    fake = _FakeFlat(m)
    graphed.FLAT_BY_MODEL[id(m)] = fake
    try:
        gen = m.parameters()
        got = FO._whole_model_generator(gen)
        assert got is not None and got[0] is fake and got[1] is m
        assert len(list(gen)) == len(list(m.parameters())) == 4              # still intact
        assert FO._whole_model_generator(m.parameters(recurse=False)) is None
        assert FO._whole_model_generator(m[0].parameters()) is None           # a sub-module's parameters are not the whole model
        other = _mlp(8)
        graphed.FLAT_BY_MODEL[id(other)] = fake                                # stale id -> weakref points elsewhere
        assert FO._whole_model_generator(other.parameters()) is None
        graphed.FLAT_BY_MODEL.pop(id(other))
    finally:
        graphed.FLAT_BY_MODEL.pop(id(m), None)
