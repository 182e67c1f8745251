"""Launcher under torchrun semantics: every rank must iterate a DISJOINT shard of the same shuffle while the global RNG
streams stay in lock-step, only rank 0 may write checkpoints / create the experiment directory without racing, and the lazily
discovered parameter registry must have the same order (and hashed names) on every rank.  World 2 on gloo, CPU."""
import os
import socket
import sys
import tempfile

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, out):
    sys.path.insert(0, ROOT)
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port), "FSB_DP_BACKEND": "gloo"})
    import torch.distributed as dist
    import torch.nn as nn
    import torch.utils.data as tud
    from fasterseg_b200 import launch, parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parallel.seed_all_ranks_identically(123)
    launch.install_rank_sharded_loaders(rank, world)
    launch.install_rank0_side_effects(rank)
    # --- what an unmodified driver does -----------------------------------------------------------------------------
    ds = tud.TensorDataset(torch.arange(40).float().view(40, 1))
    loader = tud.DataLoader(ds, batch_size=4, shuffle=True, drop_last=True)
    seen = [int(v) for (b,) in loader for v in b.view(-1)]
    seen2 = [int(v) for (b,) in loader for v in b.view(-1)]          # second epoch: a different shuffle
    draws = (float(torch.rand(1)), float(__import__("numpy").random.rand()))   # global streams still in lock-step
    exp = os.path.join(tmp, "exp")
    os.mkdir(exp)                       # create_exp_dir: both ranks call it
    os.mkdir(os.path.join(exp, "scripts"))
    torch.save({"rank": rank}, os.path.join(exp, "weights_%d.pt" % rank))
    # --- lazily discovered registry + gradient all-reduce ----------------------------------------------------------------
    torch.manual_seed(7)
    net_b = nn.Sequential(nn.Linear(3, 2), nn.Linear(2, 1))
    net_a = nn.Linear(3, 3)
    params, names = launch.discover_parameters()
    gs = parallel.GradSync(params)
    gs.set_names(names)
    x = torch.full((1, 3), float(rank + 1))
    (net_b(x).sum() + net_a(x).sum()).backward()
    gs.sync()
    g = torch.cat([p.grad.reshape(-1) for p in list(net_a.parameters()) + list(net_b.parameters())])
    import pickle
    with open(out % rank, "wb") as f:
        pickle.dump({"seen": seen, "seen2": seen2, "draws": draws, "names": names, "grad": g.tolist(), "files": sorted(os.listdir(exp))}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_launcher_shards_data_keeps_rng_lockstep_and_only_rank0_writes():
    import pickle
    tmp = tempfile.mkdtemp()
    out = os.path.join(tmp, "r%d.pkl")
    port = _free_port()
    mp.start_processes(_worker, args=(2, port, tmp, out), nprocs=2, join=True, start_method="spawn")
    r = [pickle.load(open(out % i, "rb")) for i in range(2)]
    a, b = set(r[0]["seen"]), set(r[1]["seen"])
    assert len(a) == len(b) == 20 and not (a & b) and (a | b) == set(range(40)), "ranks must see disjoint halves of one shuffle"
    assert r[0]["seen"] != r[0]["seen2"], "a new epoch reshuffles"
    assert r[0]["draws"] == r[1]["draws"], "loader shuffling must not consume the lock-step global RNG streams"
    assert r[0]["names"] == r[1]["names"] and len(r[0]["names"]) == 6
    assert r[0]["grad"] == r[1]["grad"], "all ranks hold the averaged gradient"
    files = r[0]["files"]
    assert "weights_0.pt" in files and "weights_1.pt" not in files and "scripts" in files
