"""Data-parallel training path end to end on CPU: world_size-2 `gloo`, the REAL training units (ConvBnActFn's SyncBN branch,
FactorizedReduceFn, GradSync) on the torch stand-in backend of tests/cpu_backend.py, against one process on the concatenated
batch.  SyncBN over N ranks must equal big-batch BatchNorm: forward, running statistics, dx of the local shard, and -- after
the end-of-backward gradient average -- every parameter gradient, including gamma / beta (whose SyncBN backward uses
rank-LOCAL sums exactly so that the DP average does not scale them by the world size)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PER = 2  # images per rank


def _unit_step(unit, x, t, scale):
    from fasterseg_b200 import autograd as AG
    from fasterseg_b200 import functional as F_
    for p in unit.parameters():
        p.grad = None
    for m in unit.modules():
        if isinstance(m, torch.nn.BatchNorm2d) and m.running_mean is not None:
            m.running_mean.zero_(); m.running_var.fill_(1.0)
    xh = F_.to_nhwc_half(x).detach().requires_grad_(True)
    y = unit(xh)
    y.backward(F_.to_nhwc_half(t * (scale * AG.GRAD_SCALE)))
    return (F_.to_nchw(y.detach()), F_.to_nchw(xh.grad) / AG.GRAD_SCALE,
            {k: p.grad.clone() for k, p in unit.named_parameters() if p.grad is not None},
            {k: v.clone() for k, v in unit.state_dict().items() if "running" in k})


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    torch.set_num_threads(2)
    from fasterseg_b200 import autograd as AG
    from fasterseg_b200 import engine, parallel
    from fasterseg_b200.operations import FactorizedReduce
    from fasterseg_b200.seg_oprs import ConvBnRelu
    from tests import cpu_backend
    parallel.init_from_env(backend="gloo")
    AG.set_grad_scale(16.0)
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
    try:
        with cpu_backend.installed():
            report = {}
            for name, make, cin, cout, hw_out in (
                    ("conv3x3+bn+relu", lambda: ConvBnRelu(16, 24, 3, 1, 1), 16, 24, (12, 20)),
                    ("conv3x3s2+bn+relu", lambda: ConvBnRelu(16, 24, 3, 2, 1), 16, 24, (6, 10)),
                    ("factorized_reduce", lambda: FactorizedReduce(16, 32, stride=2, slimmable=False), 16, 32, (6, 10))):
                parallel.seed_all_ranks_identically(5)
                unit = make().train()
                with torch.no_grad():
                    for m in unit.modules():
                        if isinstance(m, torch.nn.BatchNorm2d) and m.weight is not None:
                            m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.5, 0.5)
                g = torch.Generator().manual_seed(23)
                X = torch.randn(PER * world, cin, 12, 20, generator=g)
                T = torch.randn(PER * world, cout, *hw_out, generator=g)
                sl = slice(rank * PER, (rank + 1) * PER)
                engine.enable_sync_bn(True)
                sync = parallel.GradSync(list(unit.parameters())).install()
                try:
                    y_dp, dx_dp, g_dp, rs_dp = _unit_step(unit, X[sl], T[sl], 1.0 / PER)   # mean over the LOCAL shard
                finally:
                    sync.uninstall()
                engine.enable_sync_bn(False)
                y_1, dx_1, g_1, rs_1 = _unit_step(unit, X, T, 1.0 / (PER * world))         # mean over the global batch
                errs = {"y": rel(y_dp, y_1[sl]), "dx": rel(dx_dp / world, dx_1[sl])}
                errs.update({"grad:" + k: rel(g_dp[k], g_1[k]) for k in g_1})
                errs.update({"stat:" + k: rel(rs_dp[k], rs_1[k]) for k in rs_1})
                assert set(g_dp) == set(g_1)
                report[name] = errs
            q.put((rank, report))
    finally:
        dist.destroy_process_group()


def test_syncbn_units_data_parallel_equals_big_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, report in res:
        for unit, errs in report.items():
            print("rank %d %-20s %s" % (rank, unit, "  ".join("%s %.1e" % kv for kv in errs.items())))
            for k, e in errs.items():
                # fp16 storage on both sides; the only difference is the summation order of the statistics
                assert e < 2e-3, (rank, unit, k, e)
            assert any(k.endswith("bn.weight") or k.endswith(".weight") for k in errs)
