"""World-size-2 `gloo` tests of the data-parallel plumbing (no GPU): gradient all-reduce == big-batch gradient, grad-less
parameters are skipped consistently, lock-step violation is detected, SyncBN statistics hook sums across ranks."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    from fasterseg_b200 import engine, parallel
    parallel.init_from_env(backend="gloo")
    try:
        parallel.seed_all_ranks_identically(7)
        net = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))
        dead = nn.Linear(4, 4)  # never used -> grad None on every rank
        params = list(net.parameters()) + list(dead.parameters())
        full_x = torch.randn(8, 6)
        full_y = torch.randn(8, 3)
        if mode == "mismatch" and rank == 1:
            params = params[:-1]  # different participating set -> must be detected
            (dead.weight.sum()).backward()
        idx = slice(rank * 4, rank * 4 + 4)
        sync = parallel.GradSync(params).install()
        try:
            loss = ((net(full_x[idx]) - full_y[idx]) ** 2).mean()
            if mode == "mismatch":
                with pytest.raises(RuntimeError):
                    if rank == 1:
                        sync.params = list(net.parameters())[:-1]
                    loss.backward()
                q.put((rank, "detected"))
                return
            loss.backward()
        finally:
            sync.uninstall()
        ref = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))
        ref.load_state_dict(net.state_dict())
        ((ref(full_x) - full_y) ** 2).mean().backward()
        ok = all(torch.allclose(a.grad, b.grad, atol=1e-6) for a, b in zip(net.parameters(), ref.parameters()))
        ok = ok and all(p.grad is None for p in dead.parameters()) and sync.syncs == 1
        # SyncBN statistics hook
        stats = torch.full((4,), float(rank + 1))
        out = engine.dp_allreduce_stats(stats.clone())
        ok = ok and engine.dp_world_size() == 2 and torch.allclose(out, torch.full((4,), 3.0))
        a = parallel.shard_indices(10, rank, world)
        q.put((rank, bool(ok), a))
    finally:
        dist.destroy_process_group()


def _lazy_worker(rank, world, port, q):
    """`launch.install_data_parallel()` as the unmodified drivers get it: parameters are discovered lazily at the first
    backward (the script builds its model AFTER the launcher ran), gradients are averaged over the ranks."""
    sys.path.insert(0, ROOT)
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    from fasterseg_b200 import launch
    sync = launch.install_data_parallel()
    try:
        assert sync is not None and sync.params == []
        net = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))          # built after the hooks, like the drivers do
        frozen = nn.Linear(3, 3).requires_grad_(False)
        full_x, full_y = torch.randn(8, 6), torch.randn(8, 3)                      # identical on both ranks (lock-step seed)
        idx = slice(rank * 4, rank * 4 + 4)
        ((frozen(net(full_x[idx])) - full_y[idx]) ** 2).mean().backward()
        ref = nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))
        ref.load_state_dict(net.state_dict())
        # reference gradients through torch.autograd.grad: Tensor.backward is hooked process-wide
        want = torch.autograd.grad(((frozen(ref(full_x)) - full_y) ** 2).mean(), list(ref.parameters()))
        ok = all(torch.allclose(a.grad, b, atol=1e-6) for a, b in zip(net.parameters(), want))
        ok = ok and sync.syncs == 1 and len(sync.params) >= 4 and all(p.requires_grad for p in sync.params)
        q.put((rank, bool(ok)))
    finally:
        sync.uninstall()
        dist.destroy_process_group()


def test_launcher_data_parallel_hooks_discover_parameters_lazily():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_lazy_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def _run(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (0 if mode == "ok" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_grad_allreduce_equals_big_batch_and_syncbn_hook():
    res = _run("ok")
    assert res[0][1] and res[1][1]
    assert set(res[0][2]).isdisjoint(res[1][2]) and len(res[0][2]) == len(res[1][2]) == 5


def test_lockstep_violation_is_detected():
    res = _run("mismatch")
    assert [r[1] for r in res] == ["detected", "detected"]
