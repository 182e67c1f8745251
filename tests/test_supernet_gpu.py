"""GPU parity of the search supernet (fasterseg_b200.model_search.Network_Multi_Path): forward in every width-sampling
mode against golden vectors from the UNMODIFIED reference, and `_loss` + backward (the search / pretrain step of
search/train_search.py:240-250) against the CPU oracle with the two-sided fp16-storage gate."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import fasterseg_oracle as orc
from oracle import supernet_oracle as sno
from tests import helpers as H
from tests.test_boundary_cpu import _build_supernet
from tests.test_supernet_oracle import CASE, FWD, META, cfg, inputs, make_sd, oracle_loss_and_grads

pytestmark = pytest.mark.gpu


def _load(model):
    sd = make_sd()
    own = model.state_dict()
    for k, v in sd.items():
        own[k].copy_(v)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps, m.momentum = orc.BN_EPS, orc.BN_MOMENTUM
    return sd


@pytest.mark.parametrize("tag,arch_idx,mode,train,np_seed,torch_seed", FWD)
def test_supernet_forward_vs_reference_golden(tag, arch_idx, mode, train, np_seed, torch_seed):
    z = H.load_npz("supernet.npz")
    model = _build_supernet(CASE["layers"]).cuda()
    _load(model)
    model.train(train)
    x, _ = inputs()
    if np_seed is not None:
        np.random.seed(np_seed)
    if torch_seed is not None:
        torch.manual_seed(torch_seed)
    model.arch_idx, model.prun_mode = arch_idx, mode
    with torch.no_grad():
        preds = model(x.cuda())
    torch.cuda.synchronize()
    worst = 0.0
    for i, p in enumerate(preds):
        assert p.dtype == torch.float32 and p.is_contiguous()
        ref = z["%s/pred%d" % (tag, i)]
        got = p.cpu().numpy() if train else p.cpu().numpy()[:, :, ::8, ::8]
        worst = max(worst, H.rel_err(got, ref))
    print("%s: worst norm-wise rel err over the 5 logits %.3e" % (tag, worst))
    # eval mode: running statistics, well-conditioned -> fp16 tolerance; train mode: ill-conditioned BN chains (see
    # tests/test_student_gpu.py::test_student_train_step_gradients_vs_oracle for the tolerance model)
    assert worst < (5e-3 if not train else 5e-2)
    if train:
        sd = model.state_dict()
        for k in z.files:
            if k.startswith(tag + "/after:"):
                np.testing.assert_allclose(sd[k.split("after:")[1]].cpu().numpy(), z[k], rtol=2e-2, atol=2e-3)


@pytest.mark.parametrize("tag,pretrain,np_seed,torch_seed", [("loss.pretrain", True, 11, 12), ("loss.search", "some-dir", 13, 14)])
def test_supernet_loss_backward(tag, pretrain, np_seed, torch_seed):
    z = H.load_npz("supernet.npz")
    x, tgt = inputs()
    crit = nn.CrossEntropyLoss(ignore_index=255)

    def run_oracle(emulate):
        loss, sd = oracle_loss_and_grads(pretrain, np_seed, torch_seed, emulate)
        return float(loss), {k: v.grad.numpy() for k, v in sd.items() if v.requires_grad and v.grad is not None}

    l32, g32 = run_oracle(False)
    l16, g16 = run_oracle(True)
    model = _build_supernet(CASE["layers"]).cuda()
    _load(model)
    model.train(True)
    np.random.seed(np_seed)
    torch.manual_seed(torch_seed)
    loss = model._loss(x.cuda(), tgt.cuda(), pretrain)
    loss.backward()
    torch.cuda.synchronize()
    lo = float(loss.detach())
    print("%s: loss ours %.5f | fp32 oracle %.5f | fp16-emulating oracle %.5f | reference %.5f" % (tag, lo, l32, l16, float(z[tag + "/loss"][0])))
    assert abs(lo - l32) <= 1.5 * abs(l16 - l32) + 2e-3 * abs(l32)
    grads = {k: p.grad for k, p in model.named_parameters()}
    assert sorted(k for k, g in grads.items() if g is None) == sorted(k for k in grads if k not in g32)
    checked, worst, all_ours = 0, 0.0, []
    # typical deviation that fp16 storage alone causes in this (ill-conditioned, randomly initialised) supernet
    typical = float(np.median([H.rel_err(g16[k], g32[k]) for k in g32 if np.linalg.norm(g32[k]) >= 1e-10]))
    print("%s: median err(fp16-emulating oracle, fp32 oracle) over all gradients = %.3e" % (tag, typical))
    for k, g in grads.items():
        if g is None or np.linalg.norm(g32[k]) < 1e-10:
            continue
        if k.startswith("ratio_"):
            # d loss / d(width score) = <d out, out> / score: every consumer of `out` starts with conv -> BatchNorm, which makes
            # the loss invariant to the scale of `out`, so this inner product is ~0 in exact arithmetic and what any
            # implementation (fp32 included) returns is cancellation noise.  Gate it in ABSOLUTE terms against the alpha
            # gradients, which are the same kind of quantity without the cancellation.
            scale = float(np.linalg.norm(g32["alpha_1_0"]))
            a_ours = float(np.linalg.norm(g.float().cpu().numpy() - g32[k]))
            a_emu = float(np.linalg.norm(g16[k] - g32[k]))
            print("   grad %-12s |g32| %.2e  abs err ours %.2e, emulation %.2e  (|alpha_1_0 grad| %.2e)" % (
                k, float(np.linalg.norm(g32[k])), a_ours, a_emu, scale))
            assert a_ours <= 4.0 * a_emu + 0.1 * scale, k
            checked += 1
            continue
        e_ours, e_emu = H.rel_err(g.float().cpu().numpy(), g32[k]), H.rel_err(g16[k], g32[k])
        e_emu = max(e_emu, typical)  # tiny tensors (a 4x2 beta) can be lucky in one realisation
        worst = max(worst, e_ours / (e_emu + 1e-9))
        checked += 1
        all_ours.append(e_ours)
        if k.startswith(("alpha_", "beta_", "ratio_")) or checked % 80 == 0:
            print("   grad %-40s ours vs fp32 %.2e | emulation vs fp32 %.2e" % (k, e_ours, e_emu))
        # The step is chaotic: the SAME CUDA step repeated differs run to run by ~14 % (median) in these gradients, from the
        # order of fp32 atomics alone (tools/determinism_probe.py, profiles/r1_determinism_probe.log).  "ours" and "emulation"
        # are two realisations of that noise, so a single tensor -- above all one with a handful of elements -- gets a wide
        # band; the statistically stable statement is the median over all tensors, gated tightly below.
        band = 4.0 if g.numel() < 64 else 2.5
        assert e_ours <= band * e_emu + 5e-2, "%s: ours %.3e vs emulation %.3e" % (k, e_ours, e_emu)
    med_ours = float(np.median(all_ours))
    print("%s: checked %d gradients, worst err(ours)/err(emulation) %.2f, median err ours %.3e vs emulation %.3e" % (
        tag, checked, worst, med_ours, typical))
    assert med_ours <= 1.5 * typical + 1e-2
    assert checked > 300
    assert len([k for k, g in grads.items() if g is None]) == META[tag + ".no_grad_count"]


@pytest.mark.parametrize("pretrain,np_seed,torch_seed", [(True, 21, 22), ("some-dir", 23, 24)])
def test_captured_passes_match_the_eager_step(pretrain, np_seed, torch_seed, lib_option):
    """`_loss` as CUDA-graph replays of passes captured at maximum width (fasterseg_b200/graphed.py) against the eager per-unit
    `_loss` at the real sampled widths, two optimizer steps (the second one exercises the in-graph weight re-pack and the
    release of staged gradients into fresh `param.grad`).  The extra channels a captured pass computes are exact zeros, so
    the tensor-core accumulations see the same non-zero terms in the same order: the two paths agree to rounding noise of
    the few scalar operations that differ (deterministic weight gradients are switched on to make that visible)."""
    lib_option("FSB_DETERMINISTIC", 1)
    x, tgt = inputs()
    x, tgt = x.cuda(), tgt.cuda()
    models = []
    for graph in (False, None):
        m = _build_supernet(CASE["layers"]).cuda()
        _load(m)
        with torch.no_grad():
            g = torch.Generator().manual_seed(5)
            for ps in m._arch_parameters:
                for p in ps:
                    p.add_((torch.randn(p.shape, generator=g) * 0.3).cuda())
        m.train(True)
        m.__dict__["_fsb_graph_mode"] = graph
        models.append(m)
    opts = [torch.optim.SGD([p for n, p in m.named_parameters() if not n.startswith(("alpha", "beta", "ratio"))], lr=0.02,
                            momentum=0.9, weight_decay=5e-4) for m in models]
    for step in range(2):
        losses = []
        if step == 1:
            # the chain is chaotic: after one optimizer step the 1e-4 gradient differences of step 0 grow to ~10 % (same effect as
            # fp16 storage, see test_supernet_loss_backward).  To test the MECHANISM of the second step -- in-graph re-pack of the
            # updated weights, release into fresh param.grad -- both models restart it from identical weights and statistics.
            models[1].load_state_dict(models[0].state_dict())
        for m, o in zip(models, opts):
            o.zero_grad()
            for ps in m._arch_parameters:
                for p in ps:
                    p.grad = None
            np.random.seed(np_seed + step)
            torch.manual_seed(torch_seed + step)
            loss = m._loss(x, tgt, pretrain)
            loss.backward()
            losses.append(float(loss.detach()))
        torch.cuda.synchronize()
        assert models[1].__dict__.get("_fsb_graph_runner") is not None and models[1].__dict__["_fsb_graph_runner"].capture
        assert models[0].__dict__.get("_fsb_graph_runner") is None
        g0 = {k: p.grad for k, p in models[0].named_parameters()}
        g1 = {k: p.grad for k, p in models[1].named_parameters()}
        assert sorted(k for k, g in g0.items() if g is None) == sorted(k for k, g in g1.items() if g is None)
        errs = []
        for k, a in g0.items():
            if a is None or float(a.norm()) < 1e-10:
                continue
            errs.append((float((a - g1[k]).norm() / a.norm()), k))
        errs.sort(reverse=True)
        identical = sum(1 for k, a in g0.items() if a is not None and torch.equal(a, g1[k]))
        print("step %d: loss eager %.7f captured %.7f | %d gradients, %d bit-identical, median rel diff %.2e, worst %s" % (
            step, losses[0], losses[1], len(errs), identical, errs[len(errs) // 2][0], errs[:3]))
        # forward: bit-identical statistics (per-tile partial rows do not depend on the channel count) -> identical loss.
        # backward: the BatchNorm-backward reduction partitions pixels over threads by channel count, so its fp32 sums differ in
        # the last bits between the two paths; the BatchNorm chain amplifies that to ~1e-4 (measured: median 7e-5, worst 1.5e-3)
        assert abs(losses[0] - losses[1]) <= 1e-5 * abs(losses[0])
        assert errs[len(errs) // 2][0] < 1e-3
        assert max(e for e, k in errs if not k.startswith("ratio_")) < 3e-2, errs[:5]
        s0, s1 = models[0].state_dict(), models[1].state_dict()
        for k in s0:
            if "running_" in k:
                np.testing.assert_allclose(s1[k].cpu().numpy(), s0[k].cpu().numpy(), rtol=1e-4 * (1 + 10 * step), atol=1e-5, err_msg=k)
            elif k.endswith("num_batches_tracked"):
                assert int(s0[k]) == int(s1[k]), k
        for m, o in zip(models, opts):
            nn.utils.clip_grad_norm_(m.parameters(), 5)
            o.step()
