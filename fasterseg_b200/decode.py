"""Genotype decoding: architecture parameters (alpha / beta / ratio tensors of `arch_*.pt`) -> the operator, resolution
path, down-sampling and width sequence of one branch of the derived network.

Behavioural contract = the reference's decoder (train/model_seg.py:12-135), including its side effects, because the three
decodes of a constructor (last = 0, 1, 2, model_seg.py:199-201) communicate through them:
  * `betas[1]`, `betas[2]` are replaced by their softmax on EVERY decode (so the second decode sees a softmax of a softmax);
  * alpha entries are overwritten with -inf in the caller's tensors (skip ops that were ruled out).
Pinned by tests/test_decode_fuzz_cpu.py against 120 random architectures run through the unmodified reference.

Organisation (ours): a `BranchDecoder` object holds the parameter tensors and exposes the stages -- route (betas -> where
the branch steps down), widths along the route, skip pruning / op selection -- as methods over small value objects.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch

NEG_INF = -float("inf")
SKIP = 0  # index of 'skip' in genotypes.PRIMITIVES


@dataclass
class BranchGenotype:
    ops: List[int]
    path: List[int]
    downs: List[int]
    widths: List[float]

    def as_tuple(self):
        return self.ops, self.path, self.downs, self.widths


def steps_of(path: Sequence[int]) -> List[int]:
    """Per layer: 1 if the branch steps to the next coarser scale after it, else 0 (last layer: 0)."""
    deltas = list(np.diff(np.asarray(path, dtype=np.int64)))
    if any(d not in (0, 1) for d in deltas):
        raise AssertionError("a path may only keep its scale or step down by one")
    return [int(d) for d in deltas] + [0]


def path_of(steps: Sequence[int]) -> List[int]:
    """Inverse of `steps_of`: scale index per layer, starting at 0; the last entry of `steps` is not consumed."""
    taken = np.asarray([1 if s == 1 else 0 for s in steps[:-1]], dtype=np.int64)
    return [0] + [int(v) for v in np.cumsum(taken)]


def np_softmax(x):
    e = np.exp(x)
    return e / (e.sum() + np.spacing(1))


class BranchDecoder:
    def __init__(self, alphas, betas, ratios, width_mult_list, layers: int, ignore_skip: bool = False):
        self.alphas, self.betas, self.ratios = alphas, betas, ratios
        self.width_choices = width_mult_list
        self.layers = layers
        self.ignore_skip = ignore_skip

    # -- stage 1: where does the branch step down ------------------------------------------------------------
    def route(self, last: int) -> List[int]:
        """Scale index per layer for a branch that ends at scale `last` (0: 1/8, 1: 1/16, 2: 1/32)."""
        steps = [0] * self.layers
        if last == 1:
            keep_prob = self.betas[1][1:-1, 0].cpu().numpy()       # candidates: layers 1 .. len-2 of the 1/16 betas
            steps[int(np.argmax(keep_prob)) + 1] = 1
        elif last == 2:
            first, second = 0, 1                                   # fallback when no admissible pair exists
            b1 = self.betas[1][:, 0]
            b2 = self.betas[2][:, 0]
            top = 0
            for j in range(self.layers - 4):                       # second step after layer j + 2 ...
                if j - 1 <= 1:
                    continue
                joint = b1[1:j - 1] * b2[j]                        # ... first step after layer i + 1, 1 <= i < j - 1
                k = int(torch.argmax(joint))                       # first maximum, like a strict `>` scan
                if joint[k] > top:
                    top, first, second = joint[k], k + 1, j
            steps[first + 1] = 1
            steps[second + 2] = 1
        path = path_of(steps)
        assert path[-1] == last
        return path

    # -- stage 2: channel width after every layer but the last ----------------------------------------------
    def widths_along(self, path: Sequence[int]) -> List[float]:
        picked = []
        for layer in range(1, len(path)):
            scale = path[layer]
            row = self.ratios[scale][layer - max(scale, 1)]
            picked.append(self.width_choices[row.argmax()])
        return picked

    # -- stage 3: operator per layer, dropping skip connections where allowed ---------------------------------
    def _row(self, path, i):
        return self.alphas[path[i]][i - path[i]]

    def _skip_weight(self, path, i):
        return torch.softmax(self._row(path, i), dim=-1)[SKIP]

    def select_ops(self, path: Sequence[int], widths: Sequence[float]) -> BranchGenotype:
        n = len(path)
        assert n == len(widths) + 1, "len(path) %d, len(widths) %d" % (n, len(widths))
        shortest = int(np.round(n / 3.)) + path[-1] * 2
        keeps_scale = [i == n - 1 or path[i] == path[i + 1] for i in range(n)]

        droppable = []                                             # [(layer, softmax weight of skip)]
        for i in range(n):
            row = self._row(path, i)
            if self.ignore_skip:
                row[SKIP] = NEG_INF
            if row.argmax() == SKIP and keeps_scale[i]:
                droppable.append((i, self._skip_weight(path, i)))

        # every stretch between two down-sampling layers (and after the last one) must keep at least one real operator:
        # if a whole stretch is droppable skips, its weakest skip is forced to become an operator
        cuts = [i for i in range(n - 1) if path[i] < path[i + 1]]
        if cuts:
            flagged = {i for i, _ in droppable}
            for lo, hi in zip(cuts, cuts[1:] + [n]):
                stretch = range(lo + 1, hi)
                if len(stretch) > 0 and all(j in flagged for j in stretch):
                    victim, lowest = -1, 1
                    for j in stretch:
                        wgt = self._skip_weight(path, j)
                        if wgt <= lowest:
                            victim, lowest = j, wgt
                    self._row(path, victim)[SKIP] = NEG_INF

        budget = n - shortest                                      # how many layers may disappear at most
        if len(droppable) > budget:
            droppable = sorted(droppable, key=lambda item: item[1], reverse=True)[:budget]
        dropped = {i for i, _ in droppable}

        out = BranchGenotype([], [], [], [])
        for i in range(n):
            row = self._row(path, i)
            choice = row.argmax()
            if choice == SKIP:
                if i in dropped:
                    if i == n - 1:
                        out.widths = out.widths[:-1]               # a dropped final layer takes the preceding width along
                    continue
                row[SKIP] = NEG_INF                                # a skip that has to stay becomes the runner-up operator
                choice = row.argmax()
            out.ops.append(choice)
            out.path.append(path[i])
            if i < len(widths):
                out.widths.append(widths[i])
        assert len(out.path) >= shortest
        return out

    # -- all stages -----------------------------------------------------------------------------------------
    def decode(self, last: int) -> BranchGenotype:
        for s in (1, 2):
            self.betas[s] = torch.softmax(self.betas[s], dim=-1)   # in place on the caller's list, on every decode
        path = self.route(last)
        g = self.select_ops(path, self.widths_along(path))
        assert len(g.ops) == len(g.path) and len(g.path) == len(g.widths) + 1, \
            "op %d, path %d, width%d" % (len(g.ops), len(g.path), len(g.widths))
        g.downs = steps_of(g.path)
        return g


# ---- the reference's function names (train/model_seg.py:12-135), for callers that import them ------------------
def softmax(x):
    return np_softmax(x)


def path2downs(path):
    return steps_of(path)


def downs2path(downs):
    return path_of(downs)


def betas2path(betas, last, layers):
    return BranchDecoder(None, betas, None, None, layers).route(last)


def path2widths(path, ratios, width_mult_list):
    return BranchDecoder(None, None, ratios, width_mult_list, len(path)).widths_along(path)


def alphas2ops_path_width(alphas, path, widths, ignore_skip=False):
    g = BranchDecoder(alphas, None, None, None, len(path), ignore_skip).select_ops(path, widths)
    return g.ops, g.path, g.widths


def network_metas(alphas, betas, ratios, width_mult_list, layers, last, ignore_skip=False):
    return BranchDecoder(alphas, betas, ratios, width_mult_list, layers, ignore_skip).decode(last).as_tuple()
