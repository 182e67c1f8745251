"""TEST INFRASTRUCTURE ONLY -- golden vectors of the search supernet from the UNMODIFIED reference
(search/model_search.py Network_Multi_Path).  Called by oracle/make_golden.py."""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.nn as nn

from oracle import fasterseg_oracle as orc
from oracle import ref_harness as rh

WML = orc.WIDTH_MULT_LIST
CASE = {"layers": 6, "hw": (128, 256), "batch": 2, "seed": 4321}


def _np(t):
    return t.detach().cpu().numpy()


def _fill(model, seed):
    from oracle.make_golden import fill_module_from_seed
    return fill_module_from_seed(model, seed)


def make_target(batch, h, w, seed):
    rs = np.random.RandomState(seed)
    t = rs.randint(0, 19, size=(batch, h, w)).astype(np.int64)
    t[rs.uniform(size=t.shape) < 0.05] = 255
    return t


def golden_supernet(golden_dir):
    ns = rh.load_reference("search", "slimmable_ops", "operations", "seg_oprs", "genotypes", "model_search")
    Net = ns.model_search.Network_Multi_Path
    layers, (H, W), B, seed = CASE["layers"], CASE["hw"], CASE["batch"], CASE["seed"]
    crit = nn.CrossEntropyLoss(ignore_index=255)

    def fresh(train=True):
        m = Net(19, layers, crit, Fch=12, width_mult_list=WML, prun_modes=['max', 'arch_ratio'],
                stem_head_width=[(1, 1), (8. / 12, 8. / 12)])
        shapes = _fill(m, seed)
        m.train(train)
        return m, shapes

    x = orc.random_input((B, 3, H, W), seed=seed + 1)
    tgt = torch.from_numpy(make_target(B, H // 8, W // 8, seed + 2))
    rec = {}
    m, shapes = fresh()
    meta = {"case": CASE, "shapes": {k: list(v) for k, v in shapes.items()},
            "param_order": [k for k, _ in m.named_parameters()]}

    def fwd(tag, arch_idx, mode, train=True, np_seed=None, torch_seed=None):
        m, _ = fresh(train)
        if np_seed is not None:
            np.random.seed(np_seed)
        if torch_seed is not None:
            torch.manual_seed(torch_seed)
        m.arch_idx = arch_idx
        m.prun_mode = mode
        with torch.no_grad():
            preds = m(x)
        for i, p in enumerate(preds):
            rec["%s/pred%d" % (tag, i)] = _np(p).astype(np.float32) if train else _np(p)[:, :, ::8, ::8].astype(np.float32)
        if train:
            sd = m.state_dict()
            for k in ("cells.3.1._op._ops.1.bn1.bn.4.running_mean", "cells.3.1._op._ops.1.bn1.bn.4.running_var",
                      "cells.2.0._op._ops.0.bn.bn.4.running_var", "stem.%d.0.conv.1.running_mean" % arch_idx):
                if k in sd:
                    rec["%s/after:%s" % (tag, k)] = _np(sd[k]).copy()

    fwd("max.a0", 0, "max")
    fwd("min.a0", 0, "min")
    fwd("random.a1", 1, "random", np_seed=5)
    fwd("arch_ratio.a1", 1, None, torch_seed=7)
    fwd("eval.max.a0", 0, "max", train=False)

    def loss_case(tag, pretrain, np_seed, torch_seed):
        m, _ = fresh(True)
        np.random.seed(np_seed)
        torch.manual_seed(torch_seed)
        loss = m._loss(x, tgt, pretrain)
        loss.backward()
        rec[tag + "/loss"] = np.array([float(loss)], dtype=np.float64)
        grads = {k: p.grad for k, p in m.named_parameters()}
        none = sorted(k for k, g in grads.items() if g is None)
        meta[tag + ".no_grad_count"] = len(none)
        meta[tag + ".no_grad_sample"] = none[::max(1, len(none) // 40)]
        keep = [k for k in grads if k.startswith(("alpha_", "beta_", "ratio_"))]
        keep += ["stem.0.0.conv.0.weight", "stem.1.0.conv.0.weight", "cells.2.1._op._ops.3.conv1.weight",
                 "cells.2.1.downsample._ops.0.conv2.weight", "cells.4.2._op._ops.4.bn2.bn.4.weight",
                 "cells.1.0._op._ops.2.bn1.bn.0.bias", "head02.0.conv_1x1.bias", "head12.1.conv_1x1.weight",
                 "refine32.1.3.conv.0.weight", "cells.5.1._op._ops.1.conv1.weight"]
        for k in keep:
            if grads.get(k) is not None:
                g = _np(grads[k]).astype(np.float32)
                if g.ndim == 4 and g.nbytes > 200_000:   # large weight gradients: every 4th output / input channel
                    rec["%s/grad.s4:%s" % (tag, k)] = np.ascontiguousarray(g[::4, ::4])
                else:
                    rec["%s/grad:%s" % (tag, k)] = g
        # global gradient norm over the SGD parameter groups (clip_grad_norm_ input, train_search.py:249)
        sq = sum(float((g.double() ** 2).sum()) for g in grads.values() if g is not None)
        rec[tag + "/grad_norm"] = np.array([sq ** 0.5], dtype=np.float64)

    loss_case("loss.pretrain", True, 11, 12)
    loss_case("loss.search", "some-dir", 13, 14)
    np.savez_compressed(os.path.join(golden_dir, "supernet.npz"), **rec)
    with open(os.path.join(golden_dir, "supernet_meta.json"), "w") as f:
        json.dump(meta, f)
