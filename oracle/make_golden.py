"""TEST INFRASTRUCTURE ONLY -- generate `tests/golden/*` by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden            # writes tests/golden/

Every tensor-level golden vector is produced by the reference's own classes
(`search/operations.py`, `search/slimmable_ops.py`, `search/seg_oprs.py`,
`search/model_search.py`, `train/model_seg.py`) on CPU fp32 with numpy-seeded inputs and
weights, because the reference ships no tests for this path.
`tests/test_oracle_golden.py` then pins `oracle/fasterseg_oracle.py` to these files, and the
`-m gpu` tests pin the CUDA path to the oracle (and to these files directly).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import fasterseg_oracle as orc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WML = orc.WIDTH_MULT_LIST


def _np(t):
    return t.detach().cpu().numpy()


def fill_module_from_seed(module: nn.Module, seed: int, randomize_bn=True):
    """Overwrite every parameter/buffer of a reference module with the numpy-seeded synthetic values
    the oracle's `random_state_dict` would produce for the same key/shape set."""
    sd = module.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    new = orc.random_state_dict(shapes, seed=seed, randomize_bn=randomize_bn)
    seen = set()
    for k in sorted(new):  # shared cells appear under several keys ("0-0", "0-1"): first key wins
        if sd[k].data_ptr() in seen:
            continue
        seen.add(sd[k].data_ptr())
        sd[k].copy_(new[k])
    for m in module.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps, m.momentum = orc.BN_EPS, orc.BN_MOMENTUM
    return {k: tuple(s) for k, s in shapes.items()}


def golden_make_divisible(ns):
    md = ns.slimmable_ops.make_divisible
    rows = []
    for c in (8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384):
        for w in WML + [0.1, 0.25, 0.9]:
            rows.append([c, w, md(c * w)])
    for v in (0.3, 1, 3.9, 4, 7.5, 11.9, 12, 20, 100.4):
        rows.append([v, 1.0, md(v)])
    with open(os.path.join(GOLDEN, "make_divisible.json"), "w") as f:
        json.dump(rows, f)


def golden_genotypes(ns_train):
    out = {}
    for arch_idx in (0, 1):
        pristine = rh.load_arch(arch_idx)  # the decoder mutates the alpha tensors in place -> dump a fresh load
        model, state, lasts = rh.build_reference_student(ns_train, arch_idx)
        entry = {
            "arch": {k: (_np(v).tolist() if torch.is_tensor(v) else float(v)) for k, v in pristine.items()},
            "lasts": lasts,
            "decoded": {},
            "param_count_eval_build": int(sum(p.numel() for p in model.parameters())),
        }
        for last in (0, 1, 2):
            entry["decoded"][str(last)] = {
                "ops": [int(o) for o in getattr(model, "ops%d" % last)],
                "path": [int(o) for o in getattr(model, "path%d" % last)],
                "downs": [int(o) for o in getattr(model, "downs%d" % last)],
                "widths": [float(o) for o in getattr(model, "widths%d" % last)],
            }
        entry["branch_groups"] = model.branch_groups
        entry["ch_16"], entry["ch_8_2"], entry["ch_8_1"] = int(model.ch_16), int(model.ch_8_2), int(model.ch_8_1)
        entry["cells"] = {k: [int(c._C_in), int(c._C_out), int(bool(c._down)), type(c._op._op).__name__]
                          for k, c in model.cells.items()}
        entry["state_dict_shapes"] = {k: list(v.shape) for k, v in model.state_dict().items()}
        if arch_idx == 1:
            lat, size = model.forward_latency((3, 1024, 2048))
            entry["forward_latency_1024x2048"] = [float(lat), list(size)]
            m2, _, _ = rh.build_reference_student(ns_train, arch_idx, lasts=[2, 0])
            lat2, _ = m2.forward_latency((3, 1024, 2048))
            entry["forward_latency_1024x2048_lasts20"] = float(lat2)
            # train-mode build adds heads16/32 (model_seg.py:217-226)
            m3, _, _ = rh.build_reference_student(ns_train, arch_idx, train_mode=True)
            entry["state_dict_shapes_train"] = {k: list(v.shape) for k, v in m3.state_dict().items()}
        out["arch_%d" % arch_idx] = entry
    with open(os.path.join(GOLDEN, "genotypes.json"), "w") as f:
        json.dump(out, f)


def gen_x(seed, shape):
    return np.random.RandomState(seed * 3 + 1).standard_normal(shape).astype(np.float32)


def gen_gy(seed, shape):
    return np.random.RandomState(seed * 3 + 2).standard_normal(shape).astype(np.float32)


def golden_ops(ns):
    """Per-op forward (+ backward) vectors: 5 primitives + ConvNorm, stride {1,2}, non-slimmable and
    slimmable with several width pairs, eval and train mode (running-stat update included)."""
    ops = ns.operations
    cases = {}
    N, H, W = 2, 8, 12
    C_in = 24

    def run_case(name, mod, x, ratio, training, seed):
        shapes = fill_module_from_seed(mod, seed)
        if ratio is not None:
            mod.set_ratio(ratio)
        mod.train(training)
        sd_before = {k: _np(v).copy() for k, v in mod.state_dict().items() if not k.endswith("num_batches_tracked")}
        xt = torch.from_numpy(x).requires_grad_(True)
        y = mod(xt)
        rec = {"y": _np(y)}  # x / gy regenerate from the seed (gen_x / gen_gy below)
        if training:
            gy = gen_gy(seed, tuple(y.shape))
            y.backward(torch.from_numpy(gy))
            rec["gx"] = _np(xt.grad)
            for k, p in mod.named_parameters():
                if p.grad is not None:
                    rec["grad:" + k] = _np(p.grad)
            for k, v in mod.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    if not np.array_equal(_np(v), sd_before[k]):
                        rec["after:" + k] = _np(v).copy()
        for k, v in rec.items():
            cases["%s/%s" % (name, k)] = v
        meta = {"seed": seed, "ratio": ratio, "training": training, "shapes": {k: list(s) for k, s in shapes.items()}}
        return meta

    metas = {}
    seed = 1000
    builders = [
        ("BasicResidual1x", lambda ci, co, s, sl: ops.BasicResidual1x(ci, co, 3, s, 1, 1, sl, WML)),
        ("BasicResidual_downup_1x", lambda ci, co, s, sl: ops.BasicResidual_downup_1x(ci, co, 3, s, 1, 1, sl, WML)),
        ("BasicResidual2x", lambda ci, co, s, sl: ops.BasicResidual2x(ci, co, 3, s, 1, 1, sl, WML)),
        ("BasicResidual_downup_2x", lambda ci, co, s, sl: ops.BasicResidual_downup_2x(ci, co, 3, s, 1, 1, sl, WML)),
        ("FactorizedReduce", lambda ci, co, s, sl: ops.FactorizedReduce(ci, co, s, sl, WML)),
    ]
    ratio_pairs = [(1., 1.), (4. / 12, 8. / 12), (10. / 12, 6. / 12)]
    for cls_name, build in builders:
        for stride in (1, 2):
            co = C_in * stride
            for slim in (False, True):
                pairs = ratio_pairs if slim else [None]
                for pi, ratio in enumerate(pairs):
                    for training in (False, True):
                        seed += 1
                        # FactorizedReduce(stride 2, slimmable) is only self-consistent when
                        # 2*make_divisible(C_out/2*r) == make_divisible(C_out*r): true for the
                        # supernet's 96*k channels, so use C_in=48 there.
                        cin = 48 if cls_name == "FactorizedReduce" else C_in
                        co = cin * stride
                        mod = build(cin, co, stride, slim)
                        ci_act = cin if ratio is None else orc.make_divisible(cin * ratio[0])
                        x = gen_x(seed, (N, ci_act, H, W))
                        name = "%s.s%d.%s.r%d.%s" % (cls_name, stride, "slim" if slim else "fix", pi,
                                                     "train" if training else "eval")
                        if cls_name == "FactorizedReduce" and stride == 1 and not slim:
                            # identity, no parameters (operations.py:533-534)
                            y = mod(torch.from_numpy(x))
                            cases[name + "/y"] = _np(y)
                            metas[name] = {"seed": seed, "ratio": None, "training": training, "shapes": {},
                                           "cls": cls_name, "stride": stride, "slimmable": slim,
                                           "C_in": cin, "C_out": co, "x_shape": list(x.shape)}
                            continue
                        meta = run_case(name, mod, x, ratio, training, seed)
                        meta.update({"cls": cls_name, "stride": stride, "slimmable": slim, "C_in": cin, "C_out": co,
                                     "x_shape": list(x.shape)})
                        metas[name] = meta
    # odd spatial sizes for the zoomed ops (H//2, W//2 floor; operations.py:271,437)
    for cls_name, build in [builders[1], builders[3]]:
        for stride in (1, 2):
            for training in (False, True):
                seed += 1
                mod = build(C_in, C_in * stride, stride, False)
                x = gen_x(seed, (N, C_in, 9, 13))
                name = "%s.s%d.odd.%s" % (cls_name, stride, "train" if training else "eval")
                meta = run_case(name, mod, x, None, training, seed)
                meta.update({"cls": cls_name, "stride": stride, "slimmable": False, "C_in": C_in,
                             "C_out": C_in * stride, "x_shape": list(x.shape)})
                metas[name] = meta
    # ConvNorm: k in {1,3}, stride {1,2}, non-slimmable (all live uses) + one slimmable
    for k in (1, 3):
        for stride in (1, 2):
            for slim, ratio in ((False, None), (True, (1., 10. / 12))):  # forward asserts x.C == C_in (operations.py:126)
                for training in (False, True):
                    seed += 1
                    mod = ops.ConvNorm(C_in, 40, kernel_size=k, stride=stride, slimmable=slim, width_mult_list=WML)
                    ci_act = C_in if ratio is None else orc.make_divisible(C_in * ratio[0])
                    x = gen_x(seed, (N, ci_act, H, W))
                    name = "ConvNorm.k%d.s%d.%s.%s" % (k, stride, "slim" if slim else "fix", "train" if training else "eval")
                    meta = run_case(name, mod, x, ratio, training, seed)
                    meta.update({"cls": "ConvNorm", "stride": stride, "slimmable": slim, "C_in": C_in, "C_out": 40,
                                 "kernel_size": k, "x_shape": list(x.shape)})
                    metas[name] = meta
    # Head / FeatureFusion (seg_oprs.py)
    so = ns.seg_oprs
    for training in (False, True):
        seed += 1
        mod = so.Head(C_in, 19, False)
        x = gen_x(seed, (N, C_in, H, W))
        name = "Head.%s" % ("train" if training else "eval")
        meta = run_case(name, mod, x, None, training, seed)
        meta.update({"cls": "Head", "C_in": C_in, "C_out": 19, "x_shape": list(x.shape)})
        metas[name] = meta
        seed += 1
        x = gen_x(seed, (N, C_in, H, W))
        mod = so.FeatureFusion(C_in, C_in)
        name = "FeatureFusion.%s" % ("train" if training else "eval")
        # channel_attention convs exist in the state_dict but are unused (seg_oprs.py:186-195,219-225)
        meta = run_case(name, mod, x, None, training, seed)
        meta.update({"cls": "FeatureFusion", "C_in": C_in, "C_out": C_in, "x_shape": list(x.shape)})
        metas[name] = meta
    np.savez_compressed(os.path.join(GOLDEN, "ops.npz"), **cases)
    with open(os.path.join(GOLDEN, "ops_meta.json"), "w") as f:
        json.dump(metas, f)


def golden_bilinear():
    """F.interpolate(bilinear, align_corners=True) at the odd/even sizes the path uses."""
    import torch.nn.functional as F
    rs = np.random.RandomState(4242)
    cases = {}
    for i, (h, w, ho, wo) in enumerate([(12, 20, 6, 10), (6, 10, 12, 20), (7, 9, 3, 4), (3, 4, 7, 9),
                                        (8, 16, 64, 128), (5, 5, 5, 5), (1, 3, 4, 7), (9, 11, 1, 1), (2, 2, 16, 32)]):
        x = rs.standard_normal((2, 5, h, w)).astype(np.float32)
        y = F.interpolate(torch.from_numpy(x), size=(ho, wo), mode="bilinear", align_corners=True)
        cases["%d/x" % i] = x
        cases["%d/y" % i] = _np(y)
    np.savez_compressed(os.path.join(GOLDEN, "bilinear.npz"), **cases)


def golden_convnorm_gate(ns):
    """BASELINE.json configs[0]: single ConvNorm forward on a random 1x3x256x512 tensor (stem conv,
    model_seg.py:193 student 3->32; model_search.py:148 teacher 3->48)."""
    ops = ns.operations
    rec = {}
    x = orc.random_input((1, 3, 256, 512), seed=12345).numpy()
    for co in (32, 48):
        for training in (False, True):
            mod = ops.ConvNorm(3, co, kernel_size=3, stride=2, padding=1, bias=False, groups=1, slimmable=False)
            fill_module_from_seed(mod, 12345 + co)
            mod.train(training)
            y = _np(mod(torch.from_numpy(x)))
            tag = "co%d.%s" % (co, "train" if training else "eval")
            # full output is 4-6 MB fp32; keep a strided sample + moments (input/weights regenerate from seeds)
            rec[tag + "/sample"] = y[:, :, ::8, ::8].copy()
            rec[tag + "/moments"] = np.array([y.mean(), y.std(), np.abs(y).max(), (y > 0).mean()], dtype=np.float64)
            rec[tag + "/rowsum"] = y.sum(axis=(0, 1, 3)).astype(np.float64)
    np.savez_compressed(os.path.join(GOLDEN, "convnorm_gate.npz"), **rec)


def golden_student(ns_train):
    """Full derived-network forward (eval + train mode) of arch_1 (student) and arch_0 (teacher)."""
    rec = {}
    # train-mode vectors use 192x384 as well: BatchNorm over the handful of samples that a 64x128 input leaves at 1/32
    # resolution (4 per channel inside the zoomed ops) is numerically chaotic and not representative of the reference's
    # own training resolutions (>= 224x448, config_search.py:92-101)
    for arch_idx, hw in ((1, (64, 128)), (0, (64, 128)), (1, (96, 160)), (1, (192, 384))):
        for training in (False, True):
            if hw == (192, 384) and not training:
                continue
            model, state, lasts = rh.build_reference_student(ns_train, arch_idx, train_mode=training)
            fill_module_from_seed(model, 2024 + arch_idx)
            model.train(training)
            nb = 2 if training else 1
            x = orc.random_input((nb, 3) + hw, seed=99 + arch_idx)
            tag = "arch%d.%dx%d.%s" % (arch_idx, hw[0], hw[1], "train" if training else "eval")
            with torch.no_grad():
                out = model(x)
            def digest(o):
                o = _np(o).astype(np.float32)
                return o[:, :, ::4, ::4].copy(), np.array([o.mean(), o.std(), np.abs(o).max()], dtype=np.float64)

            if training:
                for name, o in zip(("pred8", "pred16", "pred32"), out):
                    rec[tag + "/" + name + ".s4"], rec[tag + "/" + name + ".moments"] = digest(o)
                sd = model.state_dict()
                for k in ("stem.0.conv.1.running_mean", "stem.0.conv.1.running_var", "heads8.conv_3x3.bn.running_var"):
                    rec[tag + "/after:" + k] = _np(sd[k]).copy()
            else:
                if arch_idx == 1 and hw == (64, 128):
                    rec[tag + "/logits"] = _np(out).astype(np.float32)
                rec[tag + "/logits.s4"], rec[tag + "/logits.moments"] = digest(out)
                rec[tag + "/argmax"] = _np(out.argmax(1)).astype(np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "student.npz"), **rec)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(12345)
    np.random.seed(12345)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ns_search = rh.load_reference("search", "slimmable_ops", "operations", "seg_oprs")
    golden_make_divisible(ns_search)
    golden_bilinear()
    golden_ops(ns_search)
    golden_convnorm_gate(ns_search)
    ns_train = rh.load_reference("train", "operations", "seg_oprs", "model_seg")
    golden_genotypes(ns_train)
    golden_student(ns_train)
    if "--supernet" in sys.argv or True:
        try:
            from oracle.make_golden_supernet import golden_supernet
        except ImportError:
            golden_supernet = None
        if golden_supernet is not None:
            golden_supernet(GOLDEN)
    sizes = {f: os.path.getsize(os.path.join(GOLDEN, f)) for f in sorted(os.listdir(GOLDEN))}
    print(json.dumps(sizes, indent=1))


if __name__ == "__main__":
    main()
