"""Data parallelism for the training steps: one process per GPU, per-image shards, ONE exchange per
optimizer step (flat-bucket gradient all-reduce) plus the per-layer BatchNorm statistics (engine.enable_sync_bn).

The reference is single-process; its drivers (search/train_search.py:215-256, train/train.py:219-271) call
`loss.backward(); clip_grad_norm_(...); optimizer.step()` and must stay unmodified, and stock DistributedDataParallel does
not fit (4 forwards per backward, ~3 100 parameters that never receive a gradient).  So the synchronisation point is
`Tensor.backward` itself: `GradSync.install()` makes every top-level backward end with an all-reduce(mean) of the
gradients of the registered parameters, before gradient clipping and the optimizer see them.

Lock-step requirement: the width sampling (`np.random.choice`, model_search.py:254-260) and the gumbel noise (`torch.rand`,
model_search.py:15) must be identical on all ranks (same sub-network everywhere, otherwise SyncBN and the gradient sets
diverge) -> `seed_all_ranks_identically()`; data sharding uses a private, rank-seeded generator (`shard_indices`).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import engine


def init_from_env(backend: Optional[str] = None, sync_bn: bool = True):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun) and enable SyncBN statistics."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    if sync_bn:
        engine.enable_sync_bn(world > 1)
    if world > 1 and os.environ.get("FSB_NATIVE_DP", "1") == "1" and torch.cuda.is_available() and dist.get_backend() == "nccl":
        init_native_dp()   # SyncBN exchange inside the fused training units, over NVLink peer memory
    return rank, local_rank, world


def init_native_dp(group=None):
    """Library-owned SyncBN exchange over NVLink peer memory (csrc/peer.cu): every rank allocates its exchange buffer, the
    64-byte CUDA IPC handles are all-gathered through torch.distributed, every rank maps its peers.  Afterwards
    `engine.dp_native()` is True: the fused training units exchange their BatchNorm statistics themselves, on the stream, with
    no host work -- also inside the captured passes of graphed.py.  (Gradients go through ONE torch.distributed / NCCL
    all-reduce of the flat staging buffer per step.)"""
    import ctypes

    from . import _lib
    lib = _lib.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    handle = (ctypes.c_char * 64)()
    _lib.check(lib.fsb_peer_alloc(handle), "fsb_peer_alloc")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    blob = b"".join(bytes(t.cpu().tolist()) for t in gathered)
    _lib.check(lib.fsb_peer_open(blob, rank, world), "fsb_peer_open")
    dist.barrier(group=group)
    engine._SYNC_BN["native"] = True
    return world


def init_native_dp_nccl(group=None):
    """(superseded by init_native_dp) give libfsb200 its own NCCL communicator for fsb_dp_allreduce_f32 on large buffers"""
    import ctypes

    from . import _lib
    lib = _lib.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (ctypes.c_char * 128)()
    if rank == 0:
        _lib.check(lib.fsb_dp_unique_id(buf), "fsb_dp_unique_id")
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    carrier = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    dist.broadcast(carrier, src=0, group=group)
    ident = bytes(carrier.cpu().tolist())
    _lib.check(lib.fsb_dp_init(ident, rank, world), "fsb_dp_init")
    engine._SYNC_BN["native"] = True
    return world


def seed_all_ranks_identically(seed: int = 12345):
    """Same torch / numpy streams on every rank (reference seed: search/config_search.py:16)."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def shard_indices(n_items: int, rank: int, world: int, epoch_seed: int = 0) -> List[int]:
    """Disjoint per-rank index shards from a PRIVATE generator (must not touch the lock-step global streams)."""
    perm = np.random.RandomState(10007 + epoch_seed).permutation(n_items)
    per = n_items // world
    return perm[rank * per:(rank + 1) * per].tolist()


class GradSync:
    """Flat-bucket all-reduce(mean) of parameter gradients.

    Parameters whose grad is None are skipped; with lock-step sampling the set is identical on every rank, which `sync()`
    verifies with a two-float handshake before the buckets go out."""

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 256 << 20, group=None):
        self.params = [p for p in params]
        self.bucket_bytes = bucket_bytes
        self.group = group
        self._orig_backward = None
        self.syncs = 0
        self.names = None      # optional parameter names (same order as params): hashed into the handshake

    def set_names(self, names):
        assert len(names) == len(self.params)
        self.names = list(names)

    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    @torch.no_grad()
    def sync(self):
        world = self.world()
        if world == 1:
            return
        live = [(i, p) for i, p in enumerate(self.params) if p.grad is not None]
        if live:
            dev = live[0][1].grad.device
        elif self.params:
            dev = self.params[0].device
        else:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        # fixed-size handshake first, ALWAYS (also with nothing live: a rank that returned early would leave the others hanging in
        # the collective): count, index sum, and a hash of the NAMES and SHAPES of the live parameters, so that a rank-dependent
        # parameter order cannot silently mix gradients of different tensors
        import zlib
        sig = "|".join("%s:%s" % (self.names[i] if self.names else i, tuple(p.shape)) for i, p in live)
        h = zlib.crc32(sig.encode())
        check = torch.tensor([float(len(live)), float(sum(i for i, _ in live) % 1000003), float(h & 0xFFFFF), float(h >> 20)],
                             device=dev, dtype=torch.float64)
        lo, hi = check.clone(), check.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        if not (torch.equal(lo, check) and torch.equal(hi, check)):
            raise RuntimeError("ranks disagree on which parameters received gradients (count / order / names / shapes) -- the "
                               "sampling RNG streams are not in lock-step (see seed_all_ranks_identically) or the parameter "
                               "registries differ")
        if not live:
            return
        buckets, cur, cur_bytes = [], [], 0
        for _, p in live:
            nbytes = p.grad.numel() * 4
            if cur and cur_bytes + nbytes > self.bucket_bytes:
                buckets.append(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            buckets.append(cur)
        for bi, bucket in enumerate(buckets):
            flat = torch.cat([p.grad.reshape(-1).float() for p in bucket])
            dist.all_reduce(flat, group=self.group)
            flat /= world
            at = 0
            for p in bucket:
                n = p.grad.numel()
                p.grad.copy_(flat[at:at + n].view_as(p.grad))
                at += n
        self.syncs += 1

    # -- hook the end of every top-level backward ------------------------------------------------
    def install(self):
        if self._orig_backward is not None:
            return self
        orig = torch.Tensor.backward
        me = self

        def backward(tensor, *args, **kwargs):
            out = orig(tensor, *args, **kwargs)
            me.sync()
            return out

        self._orig_backward = orig
        torch.Tensor.backward = backward
        return self

    def uninstall(self):
        if self._orig_backward is not None:
            torch.Tensor.backward = self._orig_backward
            self._orig_backward = None
