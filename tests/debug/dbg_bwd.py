import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests import helpers as H
from tests.test_ops_gpu import _build, _load, OPS_META, TRAIN_CASES
from fasterseg_b200 import autograd as AG, functional as F_
z = H.load_npz("ops.npz")
AG.set_grad_scale(16.0)
for name in TRAIN_CASES:
    meta = OPS_META[name]
    mod = _build(meta).cuda(); _load(mod, meta)
    if meta.get("slimmable") and meta["ratio"] is not None: mod.set_ratio(tuple(meta["ratio"]))
    mod.train(True)
    x = torch.from_numpy(H.gen_x(meta["seed"], tuple(meta["x_shape"]))).cuda()
    xh = F_.to_nhwc_half(x).detach().requires_grad_(True)
    y = mod(xh)
    ref_y = z[name + "/y"]
    yn = F_.to_nchw(y.detach(), torch.float32).cpu().numpy()
    flips = int(((yn > 0) != (ref_y > 0)).sum())
    gy = torch.from_numpy(H.gen_gy(meta["seed"], ref_y.shape)).cuda()
    y.backward(F_.to_nhwc_half(gy * AG.GRAD_SCALE))
    gx = F_.to_nchw(xh.grad, torch.float32).cpu().numpy() / AG.GRAD_SCALE
    errs = {}
    for k, p in mod.named_parameters():
        if p.grad is not None and (name + "/grad:" + k) in z.files:
            errs[k] = H.rel_err(p.grad.float().cpu().numpy(), z[name + "/grad:" + k])
    worst = max(errs.items(), key=lambda kv: kv[1])
    print("%-48s yerr %.1e flips(out) %3d gx %.1e worst %s %.1e" % (name, H.rel_err(yn, ref_y), flips, H.rel_err(gx, z[name + "/gx"]), worst[0], worst[1]))
