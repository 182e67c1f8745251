#!/usr/bin/env python
"""Run-to-run probe of the train-mode student step: the same step twice from the same state, comparing every conv+BN+act
unit's output (forward) and incoming gradient (backward) in call order.  A smooth growth of the difference with depth is the
ill-conditioned chain amplifying the order of fp32 atomics (statistics, weight gradients); a jump at one unit would be a race.
Usage: python tools/determinism_probe.py [--hw 192 384] [--batch 4]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_weights_  # noqa: E402
from fasterseg_b200 import engine, zoo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, nargs=2, default=[192, 384])
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()
    torch.manual_seed(7)
    model = zoo.build_network(1, training=True).cuda().train()
    synth_weights_(model, 3)
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(11)
    X = torch.randn(args.batch, 3, *args.hw, generator=g).cuda()
    T = [torch.randn(args.batch, 19, *args.hw, generator=g).cuda() for _ in range(3)]
    orig = engine.conv_bn_act
    rec = {}

    def probe(x, conv, bn, relu, out=None, off=(0, 0)):
        y = orig(x, conv, bn, relu, out=out, off=off)
        i = len(rec["fwd"])
        rec["fwd"].append((tuple(y.shape), y.detach().float().clone()))
        if y.requires_grad:
            y.register_hook(lambda gr, i=i: rec["bwd"].__setitem__(i, gr.detach().float().clone()))
        return y

    runs = []
    for _ in range(2):
        model.load_state_dict(state0)
        for p in model.parameters():
            p.grad = None
        rec = {"fwd": [], "bwd": {}}
        engine.conv_bn_act = probe
        outs = model(X)
        loss = sum(((o - tt) ** 2).mean() for o, tt in zip(outs, T))
        loss.backward()
        engine.conv_bn_act = orig
        torch.cuda.synchronize()
        runs.append((float(loss), rec, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
    (l0, r0, g0), (l1, r1, g1) = runs
    print("loss %.7f vs %.7f" % (l0, l1))

    def rel(a, b):
        return float((a - b).norm() / (a.norm() + 1e-30))

    print("unit  shape                      fwd rel diff   bwd(dy) rel diff")
    for i, ((shp, a), (_, b)) in enumerate(zip(r0["fwd"], r1["fwd"])):
        bw = rel(r0["bwd"][i], r1["bwd"][i]) if i in r0["bwd"] and i in r1["bwd"] else float("nan")
        print("%4d  %-26s %.3e      %.3e" % (i, str(shp), rel(a, b), bw))
    errs = sorted(((rel(g0[k], g1[k]), k) for k in g0), reverse=True)
    print("param grads: median %.3e max %.3e (%s)" % (errs[len(errs) // 2][0], errs[0][0], errs[0][1]))


if __name__ == "__main__":
    main()
