"""Graph mode of `Network_Multi_Path._loss` (fasterseg_b200/graphed.py) on the build machine: the planned passes -- every
slimmable unit at its MAXIMUM width, widths as indices in a device vector, BatchNorm sets picked from a table with the inactive
channel tail forced to zero, FactorizedReduce's width-dependent concat as a channel remap, mixing weights in static slots,
weight gradients staged in a flat buffer and released by `loss.backward()` -- must reproduce the eager `_loss` (sliced weights,
real per-width shapes), which the tests in test_hostlogic_cpu.py pin to the reference: same loss, same gradients for weights,
BatchNorm and ALL architecture parameters (alphas, betas, ratios through the straight-through gumbel sample), the same set of
grad-less parameters, the same running statistics.  Runs the passes eagerly (capture=False) on the CPU stand-in backend; the
CUDA-graph capture of exactly this code path is exercised by tests/test_supernet_gpu.py."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import fasterseg_oracle as orc
from tests import cpu_backend
from tests.test_boundary_cpu import _build_supernet
from tests.test_supernet_oracle import CASE, make_sd


def inputs():
    """eager vs graph mode is a self-comparison, so it does not need the golden case's 128x256 frames: a quarter of the pixels keeps
    every stride (down to the 2x4 map of the 1/32 branch) and the CPU suite short"""
    B, (Hh, Ww) = 2, (64, 128)
    x = orc.random_input((B, 3, Hh, Ww), seed=CASE["seed"] + 1)
    rs = np.random.RandomState(CASE["seed"] + 2)
    t = rs.randint(0, 19, size=(B, Hh // 8, Ww // 8)).astype(np.int64)
    t[rs.uniform(size=t.shape) < 0.05] = 255
    return x, torch.from_numpy(t)


@pytest.fixture(autouse=True)
def _cpu_backend():
    cpu_backend.PRECISE["on"] = True    # order-independent stand-in arithmetic: the two paths must then agree almost to the bit
    try:
        with cpu_backend.installed():
            yield
    finally:
        cpu_backend.PRECISE["on"] = False


def _model(graph):
    model = _build_supernet(CASE["layers"])
    own = model.state_dict()
    for k, v in make_sd().items():
        own[k].copy_(v)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps, m.momentum = orc.BN_EPS, orc.BN_MOMENTUM
    with torch.no_grad():     # break the symmetry of the 1e-3 * ones architecture parameters
        g = torch.Generator().manual_seed(5)
        for ps in model._arch_parameters:
            for p in ps:
                p.add_(torch.randn(p.shape, generator=g) * 0.3)
    model.train(True)
    model.__dict__["_fsb_graph_mode"] = graph
    return model


def _step(model, pretrain, np_seed, torch_seed):
    x, tgt = inputs()
    np.random.seed(np_seed)
    torch.manual_seed(torch_seed)
    loss = model._loss(x, tgt, pretrain)
    loss.backward()
    return float(loss.detach())


@pytest.mark.parametrize("pretrain,np_seed,torch_seed", [(True, 11, 12), ("some-dir", 13, 14)])
def test_graph_mode_reproduces_the_eager_loss_and_gradients(pretrain, np_seed, torch_seed):
    eager, graph = _model(False), _model(True)
    l0 = _step(eager, pretrain, np_seed, torch_seed)
    l1 = _step(graph, pretrain, np_seed, torch_seed)
    assert graph.__dict__.get("_fsb_graph_runner") is not None and eager.__dict__.get("_fsb_graph_runner") is None
    print("loss eager %.6f graph %.6f" % (l0, l1))
    assert abs(l0 - l1) <= 1e-6 * abs(l0)
    g0 = {k: p.grad for k, p in eager.named_parameters()}
    g1 = {k: p.grad for k, p in graph.named_parameters()}
    assert sorted(k for k, g in g0.items() if g is None) == sorted(k for k, g in g1.items() if g is None)
    worst = 0.0
    errs = []
    for k, a in g0.items():
        if a is None:
            continue
        b = g1[k]
        assert b.shape == a.shape and b.dtype == torch.float32
        na = float(a.norm())
        if na < 1e-10:
            assert float(b.norm()) < 1e-8, k
            continue
        e = float((a - b).norm()) / na
        errs.append((e, k))
        if k.startswith("ratio_"):     # straight-through gumbel gradients: sums of terms that cancel to ~1e-4 of their size
            assert e < 1e-2, (k, e)
            continue
        worst = max(worst, e)
    errs.sort(reverse=True)
    print("gradients compared: %d, median rel diff %.2e, worst %s" % (len(errs), errs[len(errs) // 2][0], errs[:3]))
    # with order-independent (float64-accumulating) stand-in kernels the two paths agree to fp32 rounding of a few scalar ops
    assert errs[len(errs) // 2][0] < 1e-6
    assert worst < 1e-4, errs[:5]
    for kind in ("alpha_", "beta_") + (() if pretrain is True else ("ratio_",)):
        ks = [k for k in g0 if k.startswith(kind) and g0[k] is not None]
        assert ks, kind
    s0, s1 = eager.state_dict(), graph.state_dict()
    for k in s0:
        if "running_" in k:
            np.testing.assert_allclose(s1[k].numpy(), s0[k].numpy(), rtol=2e-4, atol=1e-5, err_msg=k)
        elif k.endswith("num_batches_tracked"):
            assert int(s0[k]) == int(s1[k]), k


def test_graph_mode_second_step_sees_updated_weights_and_keeps_untouched_grads_none():
    """optimizer step between two `_loss` calls: the packed weights of the captured passes must be refreshed; parameters of
    widths that were not sampled must keep grad None so that SGD's weight decay / momentum skip them like in the reference"""
    eager, graph = _model(False), _model(True)
    opts = [torch.optim.SGD([p for n, p in m.named_parameters() if not n.startswith(("alpha", "beta", "ratio"))], lr=0.05,
                            momentum=0.9, weight_decay=5e-4) for m in (eager, graph)]
    for step in range(2):
        losses = []
        for m, o in zip((eager, graph), opts):
            o.zero_grad()
            losses.append(_step(m, True, 20 + step, 30 + step))
            nn.utils.clip_grad_norm_(m.parameters(), 5)
            o.step()
        print("step %d: eager %.6f graph %.6f" % (step, losses[0], losses[1]))
        assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[0])
    p0, p1 = dict(eager.named_parameters()), dict(graph.named_parameters())
    worst = max(float((p0[k] - p1[k]).norm() / (p0[k].norm() + 1e-12)) for k in p0)
    assert worst < 2e-3, worst   # one fp16 rounding flip of a weight after the optimizer step moves a tensor by ~1e-4
    assert {k for k, p in p0.items() if p.grad is None} == {k for k, p in p1.items() if p.grad is None}


def test_flat_buffers_and_registry_entries_go_away_with_the_model():
    """graphed.FLAT_BY_PARAM / FLAT_BY_MODEL are what the flat step tail (optim.py) looks gradients up in.  They must not keep a model
    alive: a process that builds several supernets (the reference's search script builds two, the test suite dozens) would otherwise
    accumulate 2 x 4 bytes per parameter of flat buffers per model -- and an id() of a dead parameter could be recycled."""
    import gc
    import weakref
    from fasterseg_b200 import graphed
    model = _build_supernet(3).train(True)
    model.__dict__["_fsb_graph_mode"] = True
    x = orc.random_input((2, 3, 64, 128), seed=3)
    tgt = torch.randint(0, 19, (2, 8, 16))
    np.random.seed(1)
    torch.manual_seed(2)
    model._loss(x, tgt, True).backward()
    flat = graphed.FLAT_BY_MODEL.get(id(model))
    assert flat is not None and flat.model_ref() is model
    some = next(p for p, live in zip(flat.params, flat.live_flags) if live)      # a weight whose gradient the passes staged
    assert graphed.FLAT_BY_PARAM.get(id(some)) is flat
    assert some.grad is not None and some.grad.data_ptr() == flat.gview(some).data_ptr()
    n_before = len(graphed.FLAT_BY_PARAM)
    mid, pid, fref = id(model), id(some), weakref.ref(flat)
    del flat, some, model
    gc.collect()
    assert fref() is None, "the FlatGrads (and with it ~8 bytes per parameter) outlived its model"
    assert mid not in graphed.FLAT_BY_MODEL and pid not in graphed.FLAT_BY_PARAM
    assert len(graphed.FLAT_BY_PARAM) < n_before
