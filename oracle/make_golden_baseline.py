"""TEST INFRASTRUCTURE ONLY -- golden digests of the UNMODIFIED reference at the sizes BASELINE.json's configs name
(round-1 parity only covered reduced sizes):
  C2  student arch_1 eval forward, 1 x 3 x 1024 x 2048           (train/model_seg.py:337-366)
  C3  16-layer supernet pretrain `_loss` + backward, 3 x 3 x 256 x 512   (search/model_search.py:478-505, C.pretrain=True)
  C5  16-layer supernet search `_loss` + backward, 2 x 3 x 224 x 448
Each training case also stores what the fp16-STORAGE-emulating CPU oracle (oracle/supernet_oracle.py, EMULATE_FP16) yields for
the same quantities, so that the GPU tests can state "our deviation from the fp32 reference is the deviation fp16 storage
causes" without re-running a minute-long CPU step on the GPU box.
Writes tests/golden/baseline_sizes.npz (+ .json).  Usage: python -m oracle.make_golden_baseline [c2] [c3] [c5]"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

from oracle import fasterseg_oracle as orc
from oracle import ref_harness as rh
from oracle import supernet_oracle as sno
from oracle.make_golden import fill_module_from_seed
from oracle.make_golden_supernet import make_target

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
WML = orc.WIDTH_MULT_LIST
SEED = 777
SUPERNET_CASES = {"c3": dict(pretrain=True, batch=3, hw=(256, 512), np_seed=31, torch_seed=32),
                  "c5": dict(pretrain="search-dir", batch=2, hw=(224, 448), np_seed=33, torch_seed=34)}


def _np(t):
    return t.detach().cpu().numpy()


def selected_keys(names):
    """gradient tensors stored (strided when large): all architecture parameters + a spread over depth"""
    want = [k for k in names if k.startswith(("alpha_", "beta_", "ratio_"))]
    probes = ["stem.0.0.conv.0.weight", "stem.1.0.conv.0.weight", "stem.0.2.conv2.weight", "stem.1.2.bn2.weight"]
    for layer in (0, 3, 7, 11, 15):
        for scale in (0, 1, 2):
            for tail in ("_op._ops.1.conv1.weight", "_op._ops.3.bn2.bn.4.weight", "_op._ops.4.conv2.weight",
                         "downsample._ops.0.conv2.weight", "_op._ops.0.bn.bn.2.bias", "_op._ops.2.bn1.bn.0.weight"):
                probes.append("cells.%d.%d.%s" % (layer, scale, tail))
    probes += ["refine32.0.3.conv.0.weight", "refine32.1.1.conv.1.weight", "refine16.0.1.conv.0.weight",
               "head0.0.conv_3x3.conv.weight", "head02.0.conv_1x1.bias", "head12.1.conv_1x1.weight", "head2.1.conv_3x3.bn.weight"]
    return want + [k for k in probes if k in names]


def _strided(g):
    g = _np(g).astype(np.float32)
    if g.ndim == 4 and g.nbytes > 150_000:
        g = np.ascontiguousarray(g[::4, ::4])
    if g.ndim == 4 and g.nbytes > 40_000:
        g = np.ascontiguousarray(g[::2, ::2])
    return g


def digest_grads(tag, grads, rec, meta, ref_tag=None):
    """per-tensor norms for everything; the reference's values (strided when large) for the selected tensors; for the emulating
    oracle (`ref_tag` given) only the relative deviation of every selected tensor from the reference's stored values"""
    names = [k for k, g in grads.items() if g is not None]
    meta[tag + ".grad_keys"] = names
    rec[tag + "/grad_norms"] = np.array([float(grads[k].double().norm()) for k in names], dtype=np.float64)
    rec[tag + "/grad_norm"] = np.array([float(np.sqrt((rec[tag + "/grad_norms"] ** 2).sum()))])
    sel = selected_keys(names)
    if ref_tag is None:
        meta[tag.split(".")[0] + ".selected"] = sel
        for k in sel:
            rec["%s/grad:%s" % (tag, k)] = _strided(grads[k])
    else:
        errs = []
        for k in meta[tag.split(".")[0] + ".selected"]:
            a = rec["%s/grad:%s" % (ref_tag, k)].astype(np.float64)
            b = _strided(grads[k]).astype(np.float64)
            errs.append(float(np.linalg.norm(a - b) / (np.linalg.norm(a) + 1e-30)))
        rec[tag + "/selected_rel_err"] = np.array(errs)


def supernet_case(name, rec, meta):
    c = SUPERNET_CASES[name]
    ns = rh.load_reference("search", "slimmable_ops", "operations", "seg_oprs", "genotypes", "model_search")
    Net = ns.model_search.Network_Multi_Path
    crit = nn.CrossEntropyLoss(ignore_index=255)
    B, (H, W) = c["batch"], c["hw"]
    x = orc.random_input((B, 3, H, W), seed=SEED + 1)
    tgt = torch.from_numpy(make_target(B, H // 8, W // 8, SEED + 2))
    t0 = time.time()
    m = Net(19, 16, crit, Fch=12, width_mult_list=WML, prun_modes=['max', 'arch_ratio'], stem_head_width=[(1, 1), (8. / 12, 8. / 12)])
    shapes = fill_module_from_seed(m, SEED)
    with torch.no_grad():    # break the symmetry of the architecture parameters (they start at 1e-3 * ones)
        g = torch.Generator().manual_seed(SEED + 3)
        for ps in m._arch_parameters:
            for p in ps:
                p.add_(torch.randn(p.shape, generator=g) * 0.3)
    arch = {k: _np(p).copy() for k, p in m.named_parameters() if k.startswith(("alpha_", "beta_", "ratio_"))}
    m.train(True)
    np.random.seed(c["np_seed"])
    torch.manual_seed(c["torch_seed"])
    loss = m._loss(x, tgt, c["pretrain"])
    loss.backward()
    print("%s: reference loss %.6f (%.0f s)" % (name, float(loss), time.time() - t0), flush=True)
    tag = name + ".ref"
    rec[tag + "/loss"] = np.array([float(loss)])
    grads = {k: p.grad for k, p in m.named_parameters()}
    digest_grads(tag, grads, rec, meta)
    meta[name + ".no_grad_count"] = sum(1 for g in grads.values() if g is None)
    for k, v in arch.items():
        rec["%s/arch:%s" % (name, k)] = v
    sd_after = m.state_dict()
    for k in ("cells.7.1._op._ops.1.bn1.bn.4.running_mean", "cells.7.1._op._ops.1.bn1.bn.4.running_var", "stem.0.0.conv.1.running_var"):
        rec["%s/after:%s" % (tag, k)] = _np(sd_after[k]).copy()
    del m, grads
    # the same step through the fp16-storage-emulating oracle
    t0 = time.time()
    sd = orc.random_state_dict({k: tuple(v) for k, v in shapes.items()}, seed=SEED)
    for k, v in arch.items():
        sd[k] = torch.from_numpy(v.copy())
    for k, v in sd.items():
        if "running" not in k and v.dtype.is_floating_point:
            v.requires_grad_(True)
    np.random.seed(c["np_seed"])
    torch.manual_seed(c["torch_seed"])
    orc.EMULATE_FP16["on"] = True
    try:
        loss_e = sno.supernet_loss(x, tgt, sd, sno.SupernetConfig(layers=16), crit, c["pretrain"])
        loss_e.backward()
    finally:
        orc.EMULATE_FP16["on"] = False
    print("%s: fp16-emulating oracle loss %.6f (%.0f s)" % (name, float(loss_e), time.time() - t0), flush=True)
    tag = name + ".emu"
    rec[tag + "/loss"] = np.array([float(loss_e)])
    digest_grads(tag, {k: v.grad for k, v in sd.items() if v.requires_grad}, rec, meta, ref_tag=name + ".ref")
    meta[name] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in c.items()}


def student_fullres(rec, meta):
    ns_train = rh.load_reference("train", "operations", "seg_oprs", "model_seg")
    model, state, lasts = rh.build_reference_student(ns_train, 1, train_mode=False)
    fill_module_from_seed(model, 2024 + 1)
    model.eval()
    x = orc.random_input((1, 3, 1024, 2048), seed=4242)
    t0 = time.time()
    with torch.no_grad():
        out = _np(model(x)).astype(np.float32)
    print("c2: reference forward %.1f s, logits |max| %.3f" % (time.time() - t0, np.abs(out).max()), flush=True)
    rec["c2/logits.s32"] = np.ascontiguousarray(out[:, :, 3::32, 7::32])
    rec["c2/moments"] = np.array([out.mean(), out.std(), np.abs(out).max()], dtype=np.float64)
    lab = out.argmax(1).astype(np.uint8)
    srt = np.sort(out, axis=1)
    margin = (srt[:, -1] - srt[:, -2]).astype(np.float32)
    rec["c2/argmax.s4"] = np.ascontiguousarray(lab[:, 1::4, 2::4])
    rec["c2/margin.s4"] = np.ascontiguousarray(margin[:, 1::4, 2::4]).astype(np.float16)
    rec["c2/label_hist"] = np.bincount(lab.reshape(-1), minlength=19).astype(np.int64)
    meta["c2"] = {"seed_weights": 2025, "seed_input": 4242, "hw": [1024, 2048]}


def main():
    which = [a for a in sys.argv[1:] if a in ("c2", "c3", "c5")] or ["c2", "c3", "c5"]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    path = os.path.join(GOLDEN, "baseline_sizes.npz")
    rec, meta = {}, {}
    if os.path.exists(path):
        rec = dict(np.load(path))
        meta = json.load(open(path.replace(".npz", ".json")))
    for name in which:
        for k in [k for k in rec if k.startswith(name)]:
            del rec[k]
        if name == "c2":
            student_fullres(rec, meta)
        else:
            supernet_case(name, rec, meta)
        np.savez_compressed(path, **rec)
        with open(path.replace(".npz", ".json"), "w") as f:
            json.dump(meta, f)
    print({k: os.path.getsize(os.path.join(GOLDEN, k)) for k in ("baseline_sizes.npz", "baseline_sizes.json")})


if __name__ == "__main__":
    main()
