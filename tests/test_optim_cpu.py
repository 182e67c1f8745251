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


# ---------------------------------------------------------------------------------------------------------------------------
# the flat path itself on the CPU stand-in backend (tests/cpu_backend.py implements fsb_flat_* over the same tables: raw storage
# pointers, block map, live flags), driven by the gradients of a captured-style `_loss` of a small supernet
# ---------------------------------------------------------------------------------------------------------------------------
from tests import cpu_backend  # noqa: E402
from tests.test_boundary_cpu import _build_supernet  # noqa: E402


@pytest.fixture(scope="module")
def flat_case():
    with cpu_backend.installed():
        torch.manual_seed(0)
        model = _build_supernet(3).train(True)
        model.__dict__["_fsb_graph_mode"] = True
        with torch.no_grad():
            g = torch.Generator().manual_seed(5)
            for ps in model._arch_parameters:
                for p in ps:
                    p.add_(torch.randn(p.shape, generator=g) * 0.3)
        x = torch.randn(2, 3, 64, 128, generator=torch.Generator().manual_seed(1))
        tgt = torch.randint(0, 19, (2, 8, 16), generator=torch.Generator().manual_seed(2))
        yield model, x, tgt


def _weights(model):
    return [p for n, p in model.named_parameters() if not n.startswith(("alpha", "beta", "ratio"))]


def _backward(model, x, tgt, seed, pretrain=True):
    np.random.seed(seed)
    torch.manual_seed(seed + 1)
    model._loss(x, tgt, pretrain).backward()


def test_flat_clip_and_sgd_equal_torch_arithmetic_on_the_stand_in(flat_case):
    model, x, tgt = flat_case
    ps = _weights(model)
    lr, mom, wd, max_norm = 0.05, 0.9, 5e-4, 0.5
    opt = FO.FlatSGD(ps, lr=lr, momentum=mom, weight_decay=wd)
    named = dict(model.named_parameters())
    mine = {id(p) for p in ps}
    momentum = {k: torch.zeros_like(p) for k, p in named.items()}
    for step in range(2):
        model.zero_grad(set_to_none=True)
        _backward(model, x, tgt, 10 + step, pretrain=(True if step == 0 else "search"))      # step 1: architecture parameters get gradients too
        g0 = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in named.items()}
        p0 = {k: p.detach().clone() for k, p in named.items()}
        total = torch.sqrt(sum(g.double().pow(2).sum() for g in g0.values() if g is not None)).float()
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        got = FO.clip_grad_norm_(model.parameters(), max_norm)           # the drivers' call shape: the whole-model generator
        assert float(got) == pytest.approx(float(total), rel=1e-6)
        for k, p in named.items():
            if g0[k] is not None:
                assert torch.allclose(p.grad, g0[k] * coef, rtol=1e-6, atol=1e-12), k
        opt.step()
        moved = 0
        for k, p in named.items():
            if g0[k] is None or id(p) not in mine:
                assert torch.equal(p.detach(), p0[k]), "%s has no gradient (or is not this optimizer's) but moved" % k
                continue
            momentum[k] = mom * momentum[k] + (g0[k] * coef + wd * p0[k])
            want = p0[k] - lr * momentum[k]
            assert float((p.detach() - want).abs().max()) <= 1e-6 * (float(p0[k].abs().max()) + 1e-12), (k, step)
            mb = opt.momentum_buffer(p)
            assert float((mb - momentum[k]).abs().max()) <= 1e-6 * (float(momentum[k].abs().max()) + 1e-12), (k, step)
            momentum[k] = mb.clone()
            moved += 1
        assert moved > 50
        if step == 1:
            assert any(g0[k] is not None for k in named if k.startswith(("alpha", "beta", "ratio"))), "the search step gave no arch gradient"
    assert opt.flat_steps == 2


def test_clipping_a_subset_of_the_model_takes_torchs_path(flat_case):
    model, x, tgt = flat_case
    model.zero_grad(set_to_none=True)
    _backward(model, x, tgt, 30)
    some = [p for p in _weights(model) if p.grad is not None][:7]
    g0 = [p.grad.detach().clone() for p in some]
    want = torch.sqrt(sum(g.double().pow(2).sum() for g in g0)).float()
    got = FO.clip_grad_norm_(some, 1e-3)
    assert float(got) == pytest.approx(float(want), rel=1e-5)
    coef = 1e-3 / (float(want) + 1e-6)
    for p, g in zip(some, g0):
        assert torch.allclose(p.grad, g * coef, rtol=1e-5, atol=1e-12)
    others = [p for p in _weights(model) if p.grad is not None][7:9]
    assert all(float(p.grad.abs().max()) > 0 for p in others)


def test_momentum_is_handed_to_torch_when_a_step_leaves_the_flat_path(flat_case):
    from fasterseg_b200 import graphed
    model, x, tgt = flat_case
    ps = _weights(model)
    lr, mom, wd, max_norm = 0.05, 0.9, 5e-4, 0.5
    opt = FO.FlatSGD(ps, lr=lr, momentum=mom, weight_decay=wd)
    model.zero_grad(set_to_none=True)
    _backward(model, x, tgt, 40)
    FO.clip_grad_norm_(model.parameters(), max_norm)
    opt.step()
    assert opt.flat_steps == 1
    mb = {id(p): opt.momentum_buffer(p).clone() for p in ps}
    assert sum(float(v.abs().sum()) > 0 for v in mb.values()) > 50
    model.zero_grad(set_to_none=True)
    _backward(model, x, tgt, 41)
    FO.clip_grad_norm_(model.parameters(), max_norm)
    graphed.FLAT_BY_PARAM[id(ps[0])].fresh_release = False          # e.g. gradients accumulated over two backward passes
    g0 = {id(p): p.grad.detach().clone() for p in ps if p.grad is not None}
    p0 = {id(p): p.detach().clone() for p in ps}
    opt.step()
    assert opt.flat_steps == 1, "the stale release must not take the flat path"
    checked = 0
    for p in ps:
        if id(p) not in g0:
            assert torch.equal(p.detach(), p0[id(p)])
            continue
        buf = mom * mb[id(p)] + (g0[id(p)] + wd * p0[id(p)])
        want = p0[id(p)] - lr * buf
        scale = float(p0[id(p)].abs().max()) + lr * float(buf.abs().max()) + 1e-12
        assert float((p.detach() - want).abs().max()) <= 2e-6 * scale
        assert torch.allclose(opt.momentum_buffer(p), buf, rtol=1e-5, atol=1e-9)
        checked += 1
    assert checked > 50
    # and it stays on torch's path (its state now holds the momentum), still stepping correctly
    model.zero_grad(set_to_none=True)
    _backward(model, x, tgt, 42)
    opt.step()
    assert opt.flat_steps == 1


def test_a_replaced_gradient_anywhere_in_the_model_takes_torchs_path(flat_case):
    """the flat kernels read the flat buffer, not param.grad: if ANY live parameter no longer carries the released view (here: a caller
    swapped in a hand-made gradient for one tensor deep inside the model) both calls must fall back to torch, which honours it"""
    model, x, tgt = flat_case
    ps = _weights(model)
    opt = FO.FlatSGD(ps, lr=0.1, momentum=0.0, weight_decay=0.0)
    model.zero_grad(set_to_none=True)
    _backward(model, x, tgt, 50)
    live = [p for p in ps if p.grad is not None]
    victim = live[len(live) // 2 + 3]
    victim.grad = torch.full_like(victim, 0.25)
    p0 = victim.detach().clone()
    others = {id(p): (p.detach().clone(), p.grad.detach().clone()) for p in live if p is not victim}
    assert FO._flat_of(ps) is None
    total = FO.clip_grad_norm_(model.parameters(), 1e9)      # no clipping, but the norm must include the hand-made gradient
    want = torch.sqrt(sum(p.grad.double().pow(2).sum() for p in model.parameters() if p.grad is not None)).float()
    assert float(total) == pytest.approx(float(want), rel=1e-5)
    opt.step()
    assert opt.flat_steps == 0
    assert torch.allclose(victim.detach(), p0 - 0.1 * 0.25)
    for p in live:
        if p is not victim:
            w0, g0 = others[id(p)]
            assert torch.allclose(p.detach(), w0 - 0.1 * g0, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("graph_mode", [True, False])
def test_weights_written_through_raw_pointers_invalidate_every_cache(flat_case, graph_mode, monkeypatch):
    """the flat SGD kernel writes the parameters through their storage pointers: no tensor version counter moves.  The packed-weight /
    folded-BatchNorm caches of the eager path and the weight snapshot of the captured passes key on engine.WEIGHTS_EPOCH instead; after
    a flat step the next forward -- captured or eager -- must see the NEW weights: same loss as a freshly built model holding them"""
    model, x, tgt = flat_case
    ps = _weights(model)
    opt = FO.FlatSGD(ps, lr=0.5, momentum=0.0, weight_decay=0.0)

    def loss_of(m, mode):
        m.__dict__["_fsb_graph_mode"] = mode
        m.arch_idx = 0        # like the reference (search/model_search.py:478-500) the pretrain passes run on whatever arch_idx was left behind
        np.random.seed(77)
        torch.manual_seed(78)
        return m._loss(x, tgt, True)

    try:
        model.zero_grad(set_to_none=True)
        before = loss_of(model, True)                 # gradients come from the captured-style pass (that is what fills the flat buffer)
        before.backward()
        if not graph_mode:
            with torch.no_grad():
                loss_of(model, False)                 # warm the EAGER path's caches with the old weights
        FO.clip_grad_norm_(model.parameters(), 0.5)
        # the captured passes keep their own packed copies of the weights and refresh them (the `pack` graph on a GPU) when the
        # (WEIGHTS_EPOCH, tensor versions) snapshot changed; count the refreshes as well as comparing the loss
        repacks = []
        contexts = list(model.__dict__["_fsb_graph_runner"].contexts.values())
        for c in contexts:
            monkeypatch.setattr(c, "_repack", lambda c=c, real=c._repack: (real(), repacks.append(c)))
        opt.step()
        assert opt.flat_steps == 1 and not repacks
        model.zero_grad(set_to_none=True)
        if graph_mode:                  # graph mode needs grad mode (model_search.py `_loss`); no backward: the staged gradients are dropped
            after = float(loss_of(model, True).detach())
            assert repacks and len({id(c) for c in repacks}) == len(repacks), "a captured pass kept its stale packed weights"
            repacks.clear()
            loss_of(model, True)
            assert not repacks, "weights unchanged: no refresh expected"
        else:
            with torch.no_grad():
                after = float(loss_of(model, False))
        fresh = _build_supernet(3).train(True)
        fresh.load_state_dict(model.state_dict())
        if graph_mode:
            want = float(loss_of(fresh, True).detach())
        else:
            with torch.no_grad():
                want = float(loss_of(fresh, False))
        assert abs(after - float(before.detach())) > 1e-4 * abs(want), "the step did not change the loss at all?"
        assert after == pytest.approx(want, rel=1e-5), "a cache still holds the weights from before the flat step"
    finally:
        model.__dict__["_fsb_graph_mode"] = True
