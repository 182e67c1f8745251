"""Drop-in evidence for a reference CALLER of the path: the UNMODIFIED `search/architect.py` (loaded from the reference
tree at test time, never copied) drives our `Network_Multi_Path` through the shadowed module names -- `Architect.step`
(first-order: `_loss` on the search batch, backward, Adam on the architecture parameters; with a latency weight the
expected-latency graph of `forward_latency` is back-propagated too) -- on the CPU stand-in backend.  Skipped where the
reference tree is not mounted (e.g. the GPU box)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle.make_golden_decode import SyntheticLatencyTable
from tests import cpu_backend

ARCHITECT = "/root/reference/search/architect.py"
pytestmark = pytest.mark.skipif(not os.path.isfile(ARCHITECT), reason="reference tree not mounted")


@pytest.fixture
def shadowed():
    from fasterseg_b200 import launch
    saved = {n: sys.modules.get(n) for n in launch.SHADOWED}
    launch.install_compat_patches()
    launch.install_shadow_modules()
    yield
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m


@pytest.mark.parametrize("latency_weight", [[0, 0], [0, 1e-2]])
def test_reference_architect_steps_our_supernet(shadowed, monkeypatch, latency_weight):
    from fasterseg_b200 import operations
    from fasterseg_b200.model_search import Network_Multi_Path
    monkeypatch.setattr(operations, "latency_lookup_table", SyntheticLatencyTable())
    spec = importlib.util.spec_from_file_location("ref_architect", ARCHITECT)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)                     # executes `from operations import *` against OUR operations module
    assert ref.ConvNorm is operations.ConvNorm

    with cpu_backend.installed():
        torch.manual_seed(3)
        np.random.seed(3)
        wml = [4. / 12, 6. / 12, 8. / 12, 10. / 12, 1.]
        model = Network_Multi_Path(19, 5, nn.CrossEntropyLoss(ignore_index=255), Fch=12, width_mult_list=wml,
                                   prun_modes=['max', 'arch_ratio'], stem_head_width=[(1, 1), (8. / 12, 8. / 12)]).train()
        args = types.SimpleNamespace(momentum=0.9, weight_decay=5e-4, arch_learning_rate=3e-4, latency_weight=latency_weight)
        architect = ref.Architect(model, args)
        before = [[p.detach().clone() for p in group] for group in model._arch_parameters]
        x = torch.randn(2, 3, 64, 128)
        tgt = torch.randint(0, 19, (2, 8, 16))
        loss = architect.step(x, tgt, x, tgt, unrolled=False)
        assert torch.isfinite(loss.detach()).all()
        moved = 0
        for group, old in zip(model._arch_parameters, before):
            for p, o in zip(group, old):
                assert torch.isfinite(p).all()
                if p.grad is not None and float(p.grad.abs().sum()) > 0:
                    assert not torch.equal(p.detach(), o)          # Adam moved every parameter that received a gradient
                    moved += 1
        assert moved >= 10
        if latency_weight[1] > 0:
            assert float(architect.latency_supernet.detach()) > 0  # expected latency of the student architecture (ms)
            # the latency term reaches the student's betas (they only enter through forward_latency's mixing weights
            # when a beta row has no live gradient from the loss) and its width logits
            assert model.ratio_1_0.grad is not None and float(model.ratio_1_0.grad.abs().sum()) > 0


def test_reference_init_weight_and_ohem_loss_accept_our_modules(shadowed):
    """train_search.py:77 / train.py:122 initialise the networks with the reference's `init_weight`, which picks modules by
    isinstance(nn.Conv2d / nn.BatchNorm2d) -- our slimmable classes must still qualify -- and train.py:250-258 feeds our
    logits to the reference's ProbOhemCrossEntropy2d."""
    from fasterseg_b200.model_search import Network_Multi_Path
    from fasterseg_b200.slimmable_ops import USBatchNorm2d, USConv2d
    from oracle import ref_harness
    spec = importlib.util.spec_from_file_location("ref_init_func", "/root/reference/tools/utils/init_func.py")
    init_func = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(init_func)
    wml = [4. / 12, 6. / 12, 8. / 12, 10. / 12, 1.]
    model = Network_Multi_Path(19, 5, nn.CrossEntropyLoss(ignore_index=255), Fch=12, width_mult_list=wml,
                               prun_modes=['max', 'arch_ratio'], stem_head_width=[(1, 1), (8. / 12, 8. / 12)])
    for p in model.parameters():
        nn.init.constant_(p, 7.0)
    init_func.init_weight(model, nn.init.kaiming_normal_, nn.BatchNorm2d, 1e-5, 0.1, mode='fan_in', nonlinearity='relu')
    convs = [m for m in model.modules() if isinstance(m, USConv2d)]
    bns = [m for m in model.modules() if isinstance(m, USBatchNorm2d)]
    assert len(convs) > 100 and len(bns) > 100
    assert all(float(m.weight.std()) > 0 and abs(float(m.weight.mean())) < 1.0 for m in convs)      # re-initialised
    assert all(m.eps == 1e-5 and m.momentum == 0.1 and float(m.weight.min()) == 1.0 for m in bns)   # the (unused) own affine
    assert all(float(b.weight.min()) == 1.0 and float(b.bias.abs().max()) == 0.0 for m in bns for b in m.bn)  # per-width BNs
    ns = ref_harness.load_reference("train", "seg_opr.loss_opr")
    crit = ns.modules["seg_opr.loss_opr"].ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=64, use_weight=False)
    with cpu_backend.installed():
        model.train()
        model.prun_mode, model.arch_idx = "max", 0
        logits = model(torch.randn(2, 3, 64, 128))
        loss = sum(crit(l, torch.randint(0, 19, (2, 8, 16))) for l in logits)
        loss.backward()
    assert torch.isfinite(loss.detach()) and model.stem[0][0].conv[0].weight.grad is not None
