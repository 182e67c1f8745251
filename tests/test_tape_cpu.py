"""EXPERIMENTAL tape mode (fasterseg_b200/autograd.py, FSB_TAPE=1): one torch.autograd node per network forward, our own
reverse replay inside.  On the CPU stand-in backend the taped step must reproduce the ordinary (one node per unit) step:
same loss, same set of parameters with / without gradient, the same gradients up to the rounding of a different
accumulation order of the fp16 activation gradients."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from fasterseg_b200 import autograd as AG
from oracle import fasterseg_oracle as orc
from tests import cpu_backend
from tests import helpers as H
from tests.test_boundary_cpu import _build_student, _build_supernet
from tests.test_supernet_oracle import CASE, inputs, make_sd


@pytest.fixture(autouse=True)
def _cpu_backend():
    with cpu_backend.installed():
        yield


def _supernet():
    model = _build_supernet(CASE["layers"])
    own = model.state_dict()
    for k, v in make_sd().items():
        own[k].copy_(v)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps, m.momentum = orc.BN_EPS, orc.BN_MOMENTUM
    return model.train(True)


def _step(model, pretrain, np_seed, torch_seed, taped, monkeypatch):
    monkeypatch.setattr(AG, "TAPE_ENABLED", taped)
    x, tgt = inputs()
    np.random.seed(np_seed)
    torch.manual_seed(torch_seed)
    loss = model._loss(x, tgt, pretrain)
    loss.backward()
    stats = {k: v.clone() for k, v in model.state_dict().items() if "running_mean" in k}
    return float(loss.detach()), {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()}, stats


@pytest.mark.parametrize("pretrain,np_seed,torch_seed", [(True, 11, 12), ("some-dir", 13, 14)])
def test_taped_supernet_step_equals_the_per_unit_step(monkeypatch, pretrain, np_seed, torch_seed):
    l0, g0, s0 = _step(_supernet(), pretrain, np_seed, torch_seed, False, monkeypatch)
    l1, g1, s1 = _step(_supernet(), pretrain, np_seed, torch_seed, True, monkeypatch)
    assert l1 == pytest.approx(l0, rel=1e-6)
    assert sorted(k for k, g in g0.items() if g is None) == sorted(k for k, g in g1.items() if g is None)
    errs = [H.rel_err(g1[k].numpy(), g0[k].numpy()) for k, g in g0.items() if g is not None and float(g.abs().sum()) > 0]
    print("taped vs per-unit: %d gradients, median rel diff %.2e, max %.2e" % (len(errs), float(np.median(errs)), max(errs)))
    # identical kernels and inputs; only the order in which the 10 consumers' fp16 gradients of a cell input are summed
    # differs -- and this supernet amplifies such rounding
    assert float(np.median(errs)) < 2e-2
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k          # forward is bit-identical


def test_taped_student_step_equals_the_per_unit_step(monkeypatch):
    def run(taped):
        monkeypatch.setattr(AG, "TAPE_ENABLED", taped)
        model, g = _build_student(1, training=True)
        model = model.train()
        full = {k: tuple(v) for k, v in g["state_dict_shapes_train"].items() if not k.endswith("num_batches_tracked")}
        sd = orc.random_state_dict(full, seed=31)
        own, seen = model.state_dict(), set()
        for k in sorted(sd):
            if own[k].data_ptr() not in seen:
                seen.add(own[k].data_ptr())
                own[k].copy_(sd[k])
        x = orc.random_input((2, 3, 96, 192), seed=32)
        tgt = [orc.random_input((2, 19, 96, 192), seed=33 + i) for i in range(3)]
        outs = model(x)
        assert len(outs) == 3 and all(o is not None for o in outs)
        loss = sum((o * t).mean() for o, t in zip(outs, tgt))
        loss.backward()
        return float(loss.detach()), [o.detach().clone() for o in outs], {k: (None if p.grad is None else p.grad.clone()) for k, p in model.named_parameters()}

    l0, o0, g0 = run(False)
    l1, o1, g1 = run(True)
    assert l1 == l0 and all(torch.equal(a, b) for a, b in zip(o0, o1))
    assert sorted(k for k, g in g0.items() if g is None) == sorted(k for k, g in g1.items() if g is None)
    errs = [H.rel_err(g1[k].numpy(), g0[k].numpy()) for k, g in g0.items() if g is not None and float(g.abs().sum()) > 0]
    print("student taped vs per-unit: %d gradients, median rel diff %.2e, max %.2e" % (len(errs), float(np.median(errs)), max(errs)))
    assert float(np.median(errs)) < 1e-3
