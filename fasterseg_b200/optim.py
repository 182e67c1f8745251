"""The tail of a training step on the flat gradient buffer of the captured passes (csrc/optim.cu).

The reference drivers end every step with

    nn.utils.clip_grad_norm_(model.parameters(), config.grad_clip)        # search/train_search.py:249
    optimizer.step()                                                     # search/train_search.py:250, train/train.py:269; torch.optim.SGD(lr, momentum, weight_decay)

and torch's implementations of both walk all ~5 000 Parameter objects of the supernet in Python on every step: 17 ms + 42 ms of host
time with the GPU idle, a third of a 165 ms step (profiles/r2_step_census_pretrain.log).  After a `_loss` of graph mode every weight
gradient already sits in ONE flat fp32 buffer (graphed.FlatGrads.G, `param.grad` are views into it), so both calls collapse into
three table-driven kernels: no per-parameter host work, no host synchronisation.

  * `clip_grad_norm_(parameters, max_norm)`: same return value and side effect as torch's (2-norm only); gradients that do not live in
    the flat buffer (architecture parameters after the architect step) are folded in with ordinary torch ops (a handful of tensors).
  * `FlatSGD(params, lr, momentum, weight_decay)`: `torch.optim.Optimizer` subclass with torch.optim.SGD's arithmetic (dampening 0,
    nesterov off).  It writes the parameters through their storage pointers, so it bumps `engine.WEIGHTS_EPOCH` (the packed-weight /
    folded-BN caches of the eager path key on it) instead of 5 000 tensor version counters.
  * `install()`: what the launcher does for the unmodified drivers -- `torch.optim.SGD` and `torch.nn.utils.clip_grad_norm_` are
    replaced by versions that take the flat path when it applies and fall back to torch's implementation otherwise (no flat buffer,
    gradients accumulated over several backward passes, nesterov / dampening, norm types other than 2, ...).
Momentum lives in a flat buffer too (zero-initialised: arithmetically identical to torch's "first step: buf = grad + wd * p").
"""
from __future__ import annotations

from itertools import compress

import numpy as np
import torch

from . import engine
from . import functional as F_

_TORCH_SGD = torch.optim.SGD
_TORCH_CLIP = torch.nn.utils.clip_grad_norm_


class FlatTables:
    """Device tables of one graphed.FlatGrads: segments (parameter pointer, flat offset, numel), block map, live flags."""

    def __init__(self, flat):
        self.flat = flat
        chunk = F_.flat_chunk()
        n = len(flat.params)
        seg = np.zeros(n, dtype=[("p", "<u8"), ("off", "<u4"), ("n", "<u4")])
        blocks = []
        for i, p in enumerate(flat.params):
            assert p.dtype == torch.float32 and p.is_contiguous()
            seg[i] = (p.data_ptr(), flat.offsets[id(p)], p.numel())
            nb = (p.numel() + chunk - 1) // chunk
            blocks.append(np.stack([np.full(nb, i, dtype=np.int32), np.arange(nb, dtype=np.int32)], axis=1))
        self.ptrs = seg["p"].copy()
        dev = flat.G.device
        self.segs = torch.from_numpy(seg.view(np.uint8).reshape(-1).copy()).to(dev)
        bm = np.concatenate(blocks, axis=0)
        self.nblocks = int(bm.shape[0])
        self.map = torch.from_numpy(np.ascontiguousarray(bm)).to(dev)
        self.partial = torch.empty(self.nblocks, device=dev, dtype=torch.float32)
        self.live_host = torch.zeros(n, dtype=torch.uint8).pin_memory() if dev.type == "cuda" else torch.zeros(n, dtype=torch.uint8)
        self.live = torch.zeros(n, device=dev, dtype=torch.uint8)
        self.norm = torch.zeros(2, device=dev, dtype=torch.float32)       # [total norm, clip coefficient]

    def pointers_valid(self):
        """guard against a caller that re-allocated parameter storage (p.data = ..., model.to(...)): the kernels write the weights through
        the pointers of the segment table, so EVERY pointer is compared on every use (~0.4 ms for 5 000 parameters, < 1 % of a step)"""
        cur = np.fromiter(map(torch.Tensor.data_ptr, self.flat.params), dtype=np.uint64, count=len(self.flat.params))
        return bool(np.array_equal(cur, self.ptrs))

    def set_live(self, flags: np.ndarray):
        self.live_host.numpy()[:] = flags
        self.live.copy_(self.live_host, non_blocking=True)


def _tables(flat):
    t = flat.__dict__.get("_tables")
    if t is None or not t.pointers_valid():
        t = FlatTables(flat)
        flat.__dict__["_tables"] = t
    return t


def _flat_of(params):
    """the FlatGrads whose LAST release produced the gradients of these parameters, or None"""
    from . import graphed
    for p in params:
        flat = graphed.FLAT_BY_PARAM.get(id(p))
        if flat is None:
            continue
        if flat.fresh_release is not True:
            return None
        # every live parameter must still carry the view the release handed out (not zero_grad()-ed, replaced or re-accumulated since):
        # the kernels read the flat buffer, not param.grad.  Identity of the cached view objects: ~0.8 ms for 5 000 parameters.
        views = flat._gviews
        if not all(q.grad is views.get(id(q)) for q in compress(flat.params, flat.live_flags.tolist())):
            return None
        return flat
    return None


def _whole_model_generator(parameters):
    """`model.parameters()` is a generator; consuming it walks ~5 000 modules (10 ms per step).  If the argument is exactly that
    generator for a model whose gradients live in a flat buffer, return (flat, model) WITHOUT consuming it."""
    import inspect
    from . import graphed
    if not inspect.isgenerator(parameters) or parameters.gi_code is not torch.nn.Module.parameters.__code__ or parameters.gi_frame is None:
        return None
    loc = parameters.gi_frame.f_locals
    model = loc.get("self")
    if model is None or not loc.get("recurse", True):
        return None
    flat = graphed.FLAT_BY_MODEL.get(id(model))
    return (flat, model) if flat is not None and flat.model_ref() is model else None


@torch.no_grad()
def clip_grad_norm_(parameters, max_norm, norm_type=2.0, error_if_nonfinite=False, foreach=None):
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    flat = None
    whole = _whole_model_generator(parameters) if float(norm_type) == 2.0 and not error_if_nonfinite else None
    if whole is not None:
        cand = whole[0]
        if _flat_of(cand.params[:1]) is cand:      # fresh release, gradients still the released views
            flat = cand
            live = flat.live_flags
            # gradients outside the flat buffer: parameters the captured passes never stage (architecture parameters get theirs from
            # torch autograd) -- only those that were never live can carry one
            extra = [flat.params[i].grad for i in np.nonzero(flat.ever_live == 0)[0] if flat.params[i].grad is not None]
            extra += [p.grad for p in flat.other_params if p.grad is not None]
    if flat is None:
        params = list(parameters)
        flat = _flat_of(params) if float(norm_type) == 2.0 and not error_if_nonfinite else None
        if flat is None:
            return _TORCH_CLIP(params, max_norm, norm_type=norm_type, error_if_nonfinite=error_if_nonfinite, foreach=foreach)
        live = flat.live_flags
        inside = {id(p) for p, f in zip(flat.params, live) if f}
        given = {id(p) for p in params}
        if not inside.issubset(given):      # the caller clips a subset of the model: torch's semantics need the per-tensor path
            return _TORCH_CLIP(params, max_norm, norm_type=norm_type, error_if_nonfinite=error_if_nonfinite, foreach=foreach)
        extra = [p.grad for p in params if p.grad is not None and id(p) not in inside]
    t = _tables(flat)
    t.set_live(live)
    extra_sq = None
    if extra:
        extra_sq = torch.stack([g.detach().float().pow(2).sum() for g in extra]).sum().reshape(1).contiguous()
    F_.flat_grad_norm(t.map, t.nblocks, t.segs, t.live, flat.G, t.partial, extra_sq, max_norm, t.norm)
    F_.flat_scale(t.map, t.nblocks, t.segs, t.live, flat.G, t.norm[1:])
    for g in extra:
        g.mul_(t.norm[1])
    return t.norm[0].clone()


class FlatSGD(torch.optim.Optimizer):
    """torch.optim.SGD(params, lr, momentum, weight_decay) whose step() is one kernel over the flat gradient buffer when the
    gradients of its parameters came out of a captured `_loss`, and torch's own functional SGD otherwise."""

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, **kw):
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)
        super().__init__(params, defaults)
        self._fallback = None
        self._extra = kw
        self._M = None
        self._member = None
        self.flat_steps = 0

    def _torch_sgd(self):
        if self._fallback is None:
            self._fallback = _TORCH_SGD(self.param_groups, **self._extra)      # shares the param_groups (lr schedulers keep working)
            self._fallback.param_groups = self.param_groups
        return self._fallback

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0] if len(self.param_groups) == 1 else None
        flat = _flat_of(group["params"]) if group is not None else None
        ok = flat is not None and group["dampening"] == 0 and not group["nesterov"] and not group.get("maximize", False) \
            and (self._fallback is None or not self._fallback.state)
        if not ok:
            if self._M is not None:
                self._spill_momentum()
            self._torch_sgd().step()
            return loss
        t = _tables(flat)
        if self._member is None or self._member.shape[0] != len(flat.params):
            mine = {id(p) for p in group["params"]}
            self._member = np.array([id(p) in mine for p in flat.params], dtype=np.uint8)
            self._M = torch.zeros_like(flat.G)
        live = flat.live_flags & self._member
        # every gradient this optimizer owns must be one of the released views (otherwise torch's per-tensor path is the safe one)
        t.set_live(live)
        F_.flat_sgd(t.map, t.nblocks, t.segs, t.live, flat.G, self._M, group["lr"], group["momentum"], group["weight_decay"])
        engine.bump_weights_epoch()
        self.flat_steps += 1
        return loss

    def _spill_momentum(self):
        """Leaving the flat path (gradients accumulated over several backward passes, a parameter replaced, ...): hand the momentum
        to torch's per-parameter state so that the trajectory continues; from then on this optimizer stays on torch's path."""
        from . import graphed
        fb = self._torch_sgd()
        for group in self.param_groups:
            for p in group["params"]:
                flat = graphed.FLAT_BY_PARAM.get(id(p))
                if flat is None or id(p) not in flat.offsets or self._M.numel() != flat.G.numel():
                    continue
                off = flat.offsets[id(p)]
                fb.state[p]["momentum_buffer"] = self._M[off:off + p.numel()].view(p.shape).clone()
        self._M = None
        self._member = None

    def momentum_buffer(self, p):
        """the momentum of parameter p (a view of the flat buffer), for tests / checkpoints"""
        from . import graphed
        flat = graphed.FLAT_BY_PARAM.get(id(p))
        if self._M is None or flat is None:
            st = self._torch_sgd().state.get(p, {})
            return st.get("momentum_buffer")
        off = flat.offsets[id(p)]
        return self._M[off:off + p.numel()].view(p.shape)


def _sgd_factory(params, *a, **k):
    return FlatSGD(params, *a, **k)


def install():
    """Launcher hook: `torch.optim.SGD(...)` builds a FlatSGD and `torch.nn.utils.clip_grad_norm_` is the flat-aware version."""
    torch.optim.SGD = FlatSGD
    torch.nn.utils.clip_grad_norm_ = clip_grad_norm_
    import torch.nn.utils.clip_grad as cg
    cg.clip_grad_norm_ = clip_grad_norm_


def uninstall():
    torch.optim.SGD = _TORCH_SGD
    torch.nn.utils.clip_grad_norm_ = _TORCH_CLIP
    import torch.nn.utils.clip_grad as cg
    cg.clip_grad_norm_ = _TORCH_CLIP
