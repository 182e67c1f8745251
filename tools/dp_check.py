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


def main():
    rank, local_rank, world = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
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
        errs = sorted(((float((grads_dp[k] - gref).norm() / (gref.norm() + 1e-20)), k) for k, gref in grads_1.items()), reverse=True)
        worst = errs[0][0]
        for e, k in errs[:6]:
            print("   %-60s rel diff %.3e  |g| %.3e" % (k, e, float(grads_1[k].norm())))
        print("   median rel diff %.3e" % errs[len(errs) // 2][0])
        bn_worst = max(float((bn_after[k] - v).abs().max()) for k, v in model.state_dict().items() if "running" in k)
        print("world %d: loss DP %.6f vs single-process big batch %.6f | worst grad rel diff %.3e over %d tensors | running-stat max abs diff %.2e"
              % (world, loss_dp, loss_1, worst, len(grads_1), bn_worst))
        # identical math up to fp32 summation order of the statistics / gradient reductions; the ill-conditioned train-mode
        # chain amplifies that by ~1e2..1e3
        ok = abs(loss_dp - loss_1) < 1e-4 * max(1.0, abs(loss_1)) and worst < 5e-2 and set(grads_dp) == set(grads_1)
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
