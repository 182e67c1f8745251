#!/usr/bin/env python
"""Data-parallel supernet step in graph mode (captured passes + SyncBN over NVLink peer memory + one flat gradient all-reduce)
against the single-process big batch.  Launch:  torchrun --nproc-per-node 2 tools/dp_graph_check.py
Every rank takes its shard of a global batch; afterwards rank 0 repeats the step alone on the WHOLE batch (exchange disabled)
and compares loss, every gradient and the BatchNorm running statistics.  SyncBN == big-batch BN and mean-of-rank-gradients ==
big-batch gradient hold exactly in real arithmetic; what remains is summation order (per-rank totals added in rank order vs one
pass over the big batch), amplified by the BatchNorm chain like every other last-bit difference (see tests/test_supernet_gpu.py)."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_weights_  # noqa: E402
from fasterseg_b200 import _lib, engine, parallel  # noqa: E402
from fasterseg_b200.model_search import Network_Multi_Path  # noqa: E402

WML = [4. / 12, 6. / 12, 8. / 12, 10. / 12, 1.]


def build(layers):
    m = Network_Multi_Path(19, layers, nn.CrossEntropyLoss(ignore_index=255), Fch=12, width_mult_list=WML,
                           prun_modes=['max', 'arch_ratio'], stem_head_width=[(1, 1), (8. / 12, 8. / 12)])
    synth_weights_(m, 5)
    with torch.no_grad():
        g = torch.Generator().manual_seed(6)
        for ps in m._arch_parameters:
            for p in ps:
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
    return m.cuda().train()


def step(model, x, t, pretrain):
    np.random.seed(3)
    torch.manual_seed(4)
    for p in model.parameters():
        p.grad = None
    loss = model._loss(x, t, pretrain)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach())


def main():
    layers = int(os.environ.get("LAYERS", "6"))
    rank, local_rank, world = parallel.init_from_env()
    torch.cuda.set_device(local_rank)
    _lib.set_option("FSB_DETERMINISTIC", 1)
    per = 2
    g = torch.Generator().manual_seed(99)
    X = torch.randn(per * world, 3, 128, 256, generator=g)
    T = torch.randint(0, 19, (per * world, 16, 32), generator=g)
    out = {}
    modes = {"pretrain": (True,), "search": ("dir",)}.get(os.environ.get("MODE", ""), (True, "dir"))
    for pretrain in modes:
        model = build(layers)
        xs, ts = X[rank * per:(rank + 1) * per].cuda(), T[rank * per:(rank + 1) * per].cuda()
        l_dp = step(model, xs, ts, pretrain)
        lt = torch.tensor([l_dp], device="cuda")
        torch.distributed.all_reduce(lt)
        l_dp_mean = float(lt) / world
        captured = model.__dict__.get("_fsb_graph_runner") is not None
        grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}
        stats = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k}
        # every rank must hold the same averaged gradients
        probe = torch.stack([g.double().norm() for g in grads.values() if g is not None]).sum()
        pr = [torch.zeros_like(probe) for _ in range(world)]
        torch.distributed.all_gather(pr, probe)
        same_across_ranks = all(float(abs(a - pr[0])) <= 1e-9 * float(abs(pr[0])) for a in pr)
        torch.distributed.barrier()
        if rank == 0:
            engine.enable_sync_bn(False)          # alone, on the whole batch
            big = build(layers)
            l_big = step(big, X.cuda(), T.cuda(), pretrain)
            engine.enable_sync_bn(True)
            errs = []
            none_mismatch = 0
            for k, p in big.named_parameters():
                a, b = p.grad, grads[k]
                if (a is None) != (b is None):
                    none_mismatch += 1
                    continue
                if a is None or float(a.norm()) < 1e-10:
                    continue
                errs.append((float((a - b).norm() / a.norm()), k))
            errs.sort(reverse=True)
            sd = big.state_dict()
            stat_err = max(float((sd[k] - v).norm() / (sd[k].norm() + 1e-12)) for k, v in stats.items())
            res = {"mode": "pretrain" if pretrain is True else "search", "world": world, "captured": captured,
                   "loss_dp_mean": l_dp_mean, "loss_big_batch": l_big, "grad_tensors": len(errs),
                   "grad_rel_diff_median": errs[len(errs) // 2][0], "grad_rel_diff_max": errs[0][0], "worst": errs[0][1],
                   # tensors next to the loss (heads): one unit of amplification; the median over ALL tensors includes the chain of
                   # up to 6 x 2 BatchNorm units that amplifies any last-bit difference
                   "grad_rel_diff_heads_median": (lambda h: h[len(h) // 2] if h else None)(sorted(e for e, k in errs if k.startswith("head"))),
                   "grad_none_mismatch": none_mismatch, "running_stats_rel_diff_max": stat_err,
                   "ranks_hold_identical_gradients": bool(same_across_ranks)}
            print(json.dumps(res))
            out[res["mode"]] = res
        torch.distributed.barrier()
    if rank == 0:
        ok = all(abs(r["loss_dp_mean"] - r["loss_big_batch"]) <= 1e-4 * abs(r["loss_big_batch"])
                 and (r["grad_rel_diff_heads_median"] is None or r["grad_rel_diff_heads_median"] < 5e-3) and r["grad_rel_diff_median"] < 0.3
                 and r["grad_none_mismatch"] == 0 and r["running_stats_rel_diff_max"] < 1e-3 and r["ranks_hold_identical_gradients"]
                 and r["captured"] for r in out.values())
        print("DP GRAPH CHECK", "OK" if ok else "FAILED")
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
