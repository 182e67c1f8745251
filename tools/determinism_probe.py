#!/usr/bin/env python
"""Run-to-run probe: the same training step twice from the same state, comparing every conv+BN+act unit's output (forward)
and incoming gradient (backward) in call order, the loss and every parameter gradient.

Round 1 accumulated BatchNorm statistics with fp32 atomics and the same step differed run to run by ~14 % in its
gradients (profiles/r1_determinism_probe.log).  Since round 2 every statistic is a fixed-order reduction of per-CTA partial
rows (include/fsb200.h "Deterministic statistics"): activations, activation gradients and BatchNorm parameter gradients
must be BIT-IDENTICAL; conv weight gradients still use split-K fp32 atomics by default (last-bit noise, not amplified) and
become bit-identical too with FSB_DETERMINISTIC=1 (second half of the probe).

Usage: python tools/determinism_probe.py [--hw 192 384] [--batch 4] [--supernet] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_weights_  # noqa: E402
from fasterseg_b200 import _lib, engine, zoo  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (a.double().norm() + 1e-30))


def run_twice(model, step, verbose):
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    orig = engine.conv_bn_act
    runs = []
    for _ in range(2):
        model.load_state_dict(state0)
        for p in model.parameters():
            p.grad = None
        rec = {"fwd": [], "bwd": {}}

        def probe(x, conv, bn, relu, out=None, off=(0, 0), rec=rec):
            y = orig(x, conv, bn, relu, out=out, off=off)
            i = len(rec["fwd"])
            rec["fwd"].append((tuple(y.shape), y.detach().float().clone()))
            if y.requires_grad:
                y.register_hook(lambda gr, i=i: rec["bwd"].__setitem__(i, gr.detach().float().clone()))
            return y

        engine.conv_bn_act = probe
        try:
            loss = step()
        finally:
            engine.conv_bn_act = orig
        torch.cuda.synchronize()
        runs.append((float(loss), rec, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
                     {k: v.clone() for k, v in model.state_dict().items() if "running" in k}))
    (l0, r0, g0, s0), (l1, r1, g1, s1) = runs
    fwd = [rel(a, b) for (_, a), (_, b) in zip(r0["fwd"], r1["fwd"])]
    bwd = [rel(r0["bwd"][i], r1["bwd"][i]) for i in sorted(r0["bwd"]) if i in r1["bwd"]]
    if verbose:
        print("unit  shape                      fwd rel diff   bwd(dy) rel diff")
        for i, ((shp, a), (_, b)) in enumerate(zip(r0["fwd"], r1["fwd"])):
            bw = rel(r0["bwd"][i], r1["bwd"][i]) if i in r0["bwd"] and i in r1["bwd"] else float("nan")
            print("%4d  %-26s %.3e      %.3e" % (i, str(shp), rel(a, b), bw))
    conv_w = [rel(g0[k], g1[k]) for k in g0 if g0[k].dim() == 4]
    other = [rel(g0[k], g1[k]) for k in g0 if g0[k].dim() != 4]
    stats = [rel(s0[k].float(), s1[k].float()) for k in s0]
    return {"loss": [l0, l1], "loss_identical": l0 == l1, "units": len(fwd),
            "fwd_max_rel_diff": max(fwd) if fwd else 0.0, "bwd_dy_max_rel_diff": max(bwd) if bwd else 0.0,
            "running_stats_max_rel_diff": max(stats) if stats else 0.0,
            "bn_and_scalar_grads_max_rel_diff": max(other) if other else 0.0,
            "conv_weight_grads_max_rel_diff": max(conv_w) if conv_w else 0.0,
            "conv_weight_grads_median_rel_diff": sorted(conv_w)[len(conv_w) // 2] if conv_w else 0.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, nargs=2, default=[192, 384])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--supernet", action="store_true", help="also probe a 6-layer supernet pretrain _loss")
    ap.add_argument("--json", default=None)
    ap.add_argument("-v", action="store_true")
    args = ap.parse_args()
    torch.manual_seed(7)
    model = zoo.build_network(1, training=True).cuda().train()
    synth_weights_(model, 3)
    g = torch.Generator().manual_seed(11)
    X = torch.randn(args.batch, 3, *args.hw, generator=g).cuda()
    T = [torch.randn(args.batch, 19, *args.hw, generator=g).cuda() for _ in range(3)]

    def student_step():
        outs = model(X)
        loss = sum(((o - tt) ** 2).mean() for o, tt in zip(outs, T))
        loss.backward()
        return loss

    cases = {"student": (model, student_step)}
    if args.supernet:
        import numpy as np
        import torch.nn as nn
        from fasterseg_b200.model_search import Network_Multi_Path
        widths = [4. / 12, 6. / 12, 8. / 12, 10. / 12, 1.]
        sn = Network_Multi_Path(19, 6, nn.CrossEntropyLoss(ignore_index=255), 12, widths, ['max', 'arch_ratio'],
                                [(1., 1.), (8. / 12, 8. / 12)]).cuda().train()
        synth_weights_(sn, 5)
        Xs = torch.randn(2, 3, 128, 256, generator=g).cuda()
        Ts = torch.randint(0, 19, (2, 16, 32), generator=g).cuda()

        def supernet_step():
            np.random.seed(3)          # identical width samples in both runs
            loss = sn._loss(Xs, Ts, pretrain=True)
            loss.backward()
            return loss
        cases["supernet_pretrain_6layers"] = (sn, supernet_step)

    out = {}
    for mode in (0, 1):
        _lib.set_option("FSB_DETERMINISTIC", mode)
        for name, (m, step) in cases.items():
            r = run_twice(m, step, args.v and mode == 0)
            out["%s/FSB_DETERMINISTIC=%d" % (name, mode)] = r
            print("%-40s %s" % ("%s deterministic=%d" % (name, mode), json.dumps(r)))
    _lib.set_option("FSB_DETERMINISTIC", -1)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)
    bad = [k for k, r in out.items() if r["fwd_max_rel_diff"] or r["bwd_dy_max_rel_diff"] or r["running_stats_max_rel_diff"]
           or r["bn_and_scalar_grads_max_rel_diff"] or (k.endswith("=1") and r["conv_weight_grads_max_rel_diff"])]
    print("DETERMINISM", "OK" if not bad else "FAILED: %s" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
