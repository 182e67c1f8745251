"""Run an UNMODIFIED reference driver (search/train_search.py, train/train.py, latency/run_latency.py) on the B200 path.

    cd <FasterSeg>/search && python -m fasterseg_b200.launch train_search.py
    torchrun --nproc-per-node 8 -m fasterseg_b200.launch train_search.py        # data parallel

The reference resolves `operations`, `slimmable_ops`, `seg_oprs`, `model_seg`, `model_search`, `genotypes` from cwd
(search/model_search.py:4-9, train/train.py:32); registering our modules under those names in `sys.modules` BEFORE the
script imports them swaps the whole L1/L2 layer with zero script edits.  The shim also applies the
py3.12 / torch 2.x compat patches the old scripts need (SURVEY section 5 "compat list") and, under torchrun, the
data-parallel hooks of `fasterseg_b200.parallel`.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types

SHADOWED = ("genotypes", "slimmable_ops", "operations", "seg_oprs", "model_seg", "model_search")


def install_shadow_modules():
    for name in SHADOWED:
        sys.modules[name] = importlib.import_module("fasterseg_b200." + name)
    return [sys.modules[n] for n in SHADOWED]


def install_compat_patches():
    import numpy as np
    import torch
    # `dataloader.next()` (train_search.py:226,235; train.py:237)
    from torch.utils.data.dataloader import _BaseDataLoaderIter
    if not hasattr(_BaseDataLoaderIter, "next"):
        _BaseDataLoaderIter.next = _BaseDataLoaderIter.__next__
    # pickled dicts / genotype files
    if not getattr(np.load, "_fsb_patched", False):
        orig = np.load

        def load(*a, **k):
            k.setdefault("allow_pickle", True)
            return orig(*a, **k)

        load._fsb_patched = True
        np.load = load
    if not getattr(torch.load, "_fsb_patched", False):
        torig = torch.load

        def tload(*a, **k):
            k.setdefault("weights_only", False)
            return torig(*a, **k)

        tload._fsb_patched = True
        torch.load = tload
    import collections
    import collections.abc
    for n in ("Iterable", "Mapping", "Sequence"):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    for name, attrs in (("thop", {"profile": lambda *a, **k: (0, 0)}),):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                m = types.ModuleType(name)
                m.__dict__.update(attrs)
                sys.modules[name] = m


def install_data_parallel():
    """Under torchrun: NCCL process group, SyncBN statistics, gradient all-reduce at the end of every backward for all
    parameters of every nn.Module the script builds (registered lazily at the first backward)."""
    from . import parallel
    rank, local_rank, world = parallel.init_from_env()
    if world == 1:
        return None
    import torch
    parallel.seed_all_ranks_identically()

    class _Lazy(parallel.GradSync):
        def sync(self):
            if not self.params:
                import gc
                seen = set()
                for obj in gc.get_objects():
                    if isinstance(obj, torch.nn.Module):
                        for p in obj.parameters():
                            if id(p) not in seen and p.requires_grad:
                                seen.add(id(p))
                                self.params.append(p)
            super().sync()

    return _Lazy([]).install()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m fasterseg_b200.launch <reference_script.py> [args...]")
    from . import _lib
    _lib.lib()  # fail loudly up front if the native library has not been built
    install_compat_patches()
    install_shadow_modules()
    install_data_parallel()
    script = argv[0]
    sys.argv = argv
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)) or ".")
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
