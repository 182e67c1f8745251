#!/usr/bin/env python
"""Multi-GPU data-parallel correctness: N ranks x per-rank shard == one process on the concatenated batch.
Run:  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
Every rank builds the student (train mode), takes its shard of a fixed global batch, runs forward + backward with SyncBN
statistics (engine.enable_sync_bn) and the end-of-backward gradient all-reduce (parallel.GradSync); rank 0 then repeats the
step alone on the whole batch with both switched off and compares loss and every parameter gradient."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_weights_  # noqa: E402
from fasterseg_b200 import engine, parallel, zoo  # noqa: E402


def run(model, x, t):
    for p in model.parameters():
        p.grad = None
    outs = model(x)
    loss = sum(((o - tt) ** 2).mean() for o, tt in zip(outs, t))
    loss.backward()
    return float(loss.detach()), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}


def unit_check(rank, world):
    """One conv+BN+ReLU unit (no chain, so no chaotic amplification): SyncBN forward, backward (dx of the local shard, weight /
    gamma / beta gradients after the DP average) against one process on the whole batch.  Tight tolerance (fp16 storage)."""
    from fasterseg_b200.seg_oprs import ConvBnRelu
    torch.manual_seed(5)
    unit = ConvBnRelu(32, 48, 3, 1, 1).cuda().train()
    with torch.no_grad():
        unit.bn.weight.uniform_(0.5, 1.5)
        unit.bn.bias.uniform_(-0.5, 0.5)
    per = 2
    g = torch.Generator().manual_seed(23)
    X = torch.randn(per * world, 32, 40, 72, generator=g).cuda()
    T = torch.randn(per * world, 48, 40, 72, generator=g).cuda()
    sl = slice(rank * per, (rank + 1) * per)

    from fasterseg_b200 import autograd as AG
    from fasterseg_b200 import functional as F_
    old_scale = AG.GRAD_SCALE
    AG.set_grad_scale(16.0)  # t ~ N(0,1) is orders of magnitude larger than real loss gradients

    def step(x, t, scale):
        for p in unit.parameters():
            p.grad = None
        unit.bn.running_mean.zero_(); unit.bn.running_var.fill_(1.0)
        xh = F_.to_nhwc_half(x).detach().requires_grad_(True)
        y = unit(xh)
        y.backward(F_.to_nhwc_half(t * (scale * AG.GRAD_SCALE)))   # d/dy of scale * sum(y * t), in the fp16 gradient domain
        gx = F_.to_nchw(xh.grad, torch.float32) / AG.GRAD_SCALE
        return F_.to_nchw(y.detach(), torch.float32), gx, {k: p.grad.clone() for k, p in unit.named_parameters()}

    engine.enable_sync_bn(True)
    sync = parallel.GradSync(list(unit.parameters())).install()
    # every rank back-propagates the mean over ITS shard; the DP average turns that into the global mean
    y_dp, dx_dp, g_dp = step(X[sl], T[sl], 1.0 / per)
    sync.uninstall()
    rm_dp = unit.bn.running_mean.clone()
    ok = True
    engine.enable_sync_bn(False)
    y_1, dx_1, g_1 = step(X, T, 1.0 / (per * world))
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    # dx: the DP run differentiates sum_r L_r / per, the single run sum L / (per*world): same per-sample weights up to 1/world
    errs = {"y": rel(y_dp, y_1[sl]), "dx": rel(dx_dp / world, dx_1[sl]), "running_mean": rel(rm_dp, unit.bn.running_mean)}
    errs.update({k: rel(g_dp[k], g_1[k]) for k in g_1})
    bad = {k: v for k, v in errs.items() if not v < 3e-3}
    print("[rank %d] unit check (conv3x3 32->48 + SyncBN + ReLU): %s" % (rank, "  ".join("%s %.2e" % kv for kv in errs.items())))
    engine.enable_sync_bn(True)
    AG.set_grad_scale(old_scale)
    return not bad


def main():
    rank, local_rank, world = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    unit_ok = unit_check(rank, world)
    parallel.seed_all_ranks_identically(7)
    model = zoo.build_network(1, training=True).cuda().train()
    synth_weights_(model, 3)
    state0 = {k: v.clone() for k, v in model.state_dict().items()}
    per = 2
    g = torch.Generator().manual_seed(11)
    X = torch.randn(per * world, 3, 192, 384, generator=g)
    T = [torch.randn(per * world, 19, 192, 384, generator=g) for _ in range(3)]
    sl = slice(rank * per, (rank + 1) * per)
    sync = parallel.GradSync(list(model.parameters())).install()
    # the distributed loss is the mean over the GLOBAL batch: each rank's mean over its shard, averaged by the all-reduce
    loss_dp, grads_dp = run(model, X[sl].cuda(), [t[sl].cuda() for t in T])
    sync.uninstall()
    lt = torch.tensor([loss_dp], device="cuda")
    torch.distributed.all_reduce(lt)
    loss_dp = float(lt) / world
    bn_after = {k: v.clone() for k, v in model.state_dict().items() if "running" in k}
    ok = True
    if rank == 0:
        engine.enable_sync_bn(False)
        model.load_state_dict(state0)
        loss_1, grads_1 = run(model, X.cuda(), [t.cuda() for t in T])
        # run-to-run floor: the same single-process step twice differs only by the order of the fp32 atomics in the statistics /
        # weight-gradient reductions -- the same class of perturbation the rank split introduces
        model.load_state_dict(state0)
        _, grads_2 = run(model, X.cuda(), [t.cuda() for t in T])
        floor = sorted(float((grads_2[k] - gref).norm() / (gref.norm() + 1e-20)) for k, gref in grads_1.items())
        print("   single-process run-to-run floor: median %.3e  max %.3e" % (floor[len(floor) // 2], floor[-1]))
        errs = sorted(((float((grads_dp[k] - gref).norm() / (gref.norm() + 1e-20)), k) for k, gref in grads_1.items()), reverse=True)
        worst = errs[0][0]
        for e, k in errs[:6] + errs[-6:]:
            print("   %-60s rel diff %.3e  |g| %.3e" % (k, e, float(grads_1[k].norm())))
        print("   median rel diff %.3e" % errs[len(errs) // 2][0])
        bn_worst = max(float((bn_after[k] - v).abs().max()) for k, v in model.state_dict().items() if "running" in k)
        print("world %d: loss DP %.6f vs single-process big batch %.6f | worst grad rel diff %.3e over %d tensors | running-stat max abs diff %.2e"
              % (world, loss_dp, loss_1, worst, len(grads_1), bn_worst))
        # identical math up to fp32 summation order of the statistics / gradient reductions; the ill-conditioned train-mode
        # chain amplifies that by ~1e2..1e3
        # -> gate the network-level comparison against the measured run-to-run floor, and tightly on the tensors with a short
        #    backward chain (the heads' last convs), which the chain does not amplify
        med = errs[len(errs) // 2][0]
        short = max(e for e, k in errs if ".conv_1x1." in k and k.startswith("heads"))
        print("   heads*.conv_1x1 worst rel diff %.3e" % short)
        ok = (abs(loss_dp - loss_1) < 1e-3 * max(1.0, abs(loss_1)) and set(grads_dp) == set(grads_1) and short < 1e-2
              and med < 2 * floor[len(floor) // 2] + 1e-3 and worst < 2 * floor[-1] + 1e-3)
    ok = ok and unit_ok
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    torch.distributed.broadcast(flag, 0)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    if float(flag) != 1.0:
        raise SystemExit("DP check FAILED")
    if rank == 0:
        print("DP check OK")


if __name__ == "__main__":
    main()
