#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- fuzz goldens for the genotype decoder and the derived-network builder.

Runs the UNMODIFIED reference (`train/model_seg.py:12-296`: network_metas / alphas2ops_path_width / betas2path /
path2widths / get_branch_groups_cells / build_arm_ffm_head) on randomly drawn architecture parameters -- far more decision
paths (skip pruning, down-sample placement, width picks, branch sharing) than the two shipped genotypes exercise -- and
records, per case: the decoded (ops, path, downs, widths) for last = 0, 1, 2 in constructor order (the decoder mutates
alphas / betas between calls), and for several `lasts` choices the branch groups, the per-cell (C_in, C_out, down, op class),
ch_16 / ch_8_2 / ch_8_1 and a digest of the state_dict (names + shapes); or the exception type when the reference rejects
the case.  Written to tests/golden/decode_fuzz.json; inputs are regenerated from the per-case seed by `draw_case`.
Run in the build container:  python oracle/make_golden_decode.py"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness  # noqa: E402

WML = [4. / 12, 6. / 12, 8. / 12, 10. / 12, 1.]
LASTS_CHOICES = ([2, 1], [2, 0], [1, 0], [0, 1, 2], [2], [1], [0], [1, 2])


def draw_case(seed):
    """Architecture parameters of one fuzz case (numpy MT19937: stable across platforms).  Returns plain dict of
    float32 tensors + scalar settings; called by the generator AND by the tests."""
    rs = np.random.RandomState(seed)
    layers = int(rs.choice([6, 7, 8, 9, 12, 16]))
    ignore_skip = bool(rs.rand() < 0.3)
    single_width = bool(rs.rand() < 0.15)
    nw = 1 if single_width else len(WML)
    style = rs.randint(3)  # 0: gaussian, 1: skip-heavy alphas, 2: near-uniform (ties broken by tiny noise)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))

    def alpha(rows):
        a = rs.standard_normal((rows, 5))
        if style == 1:
            a[:, 0] += rs.uniform(0.5, 3.0)
        if style == 2:
            a = 1e-3 + 1e-4 * a
        return t(a)

    case = {"layers": layers, "ignore_skip": ignore_skip, "stem_head_width": (1., 1.) if rs.rand() < 0.5 else (8. / 12, 8. / 12),
            "alphas": [alpha(layers), alpha(layers - 1), alpha(layers - 2)],
            "betas": [None, t(rs.standard_normal((layers - 2, 2))), t(rs.standard_normal((layers - 3, 2)))],
            "ratios": [t(rs.standard_normal((layers - 1, nw))), t(rs.standard_normal((layers - 1, nw))),
                       t(rs.standard_normal((layers - 2, nw)))]}
    return case


def clone_params(case):
    c = lambda ts: [None if x is None else x.clone() for x in ts]
    return c(case["alphas"]), c(case["betas"]), c(case["ratios"])


class SyntheticLatencyTable(dict):
    """Stand-in for latency_lookup_table.npy (reference DATA, not shipped here): every key is "present" and maps to a
    deterministic pseudo-random latency, so `forward_latency` is a pure function of the keys it builds and the order it
    sums them in.  Installed into the reference's modules by the generator and into ours by the tests."""

    def __contains__(self, key):
        return True

    def __getitem__(self, key):
        import zlib
        return 0.05 + (zlib.crc32(str(key).encode()) % 100003) / 100003.0


def digest(model):
    h = hashlib.sha256()
    for k, v in model.state_dict().items():
        h.update(("%s:%s;" % (k, tuple(v.shape))).encode())
    return h.hexdigest()[:16]


def describe(model):
    cells = ";".join("%s:%d,%d,%d,%s" % (k, c._C_in, c._C_out, int(bool(c._down)), type(c._op._op).__name__)
                     for k, c in sorted(model.cells.items()))
    return {"branch_groups": model.branch_groups, "cells": hashlib.sha256(cells.encode()).hexdigest()[:16],
            "ch": [model.ch_16, model.ch_8_2, model.ch_8_1],
            "digest": digest(model), "params": int(sum(p.numel() for p in model.parameters())),
            "latency_1024x2048": float(model.forward_latency((3, 1024, 2048))[0])}


def run_case(Net, case, training):
    """-> golden record for one case; `Net` is Network_Multi_Path_Infer (reference or ours)."""
    rec = {}
    alphas, betas, ratios = clone_params(case)
    try:
        model = Net(alphas, betas, ratios, num_classes=19, layers=case["layers"], Fch=12, width_mult_list=WML,
                    stem_head_width=case["stem_head_width"], ignore_skip=case["ignore_skip"])
    except Exception as e:  # noqa: BLE001 -- the reference rejects some random genotypes with assert / IndexError
        return {"ctor_error": type(e).__name__}
    rec["decoded"] = {str(last): {"ops": [int(o) for o in getattr(model, "ops%d" % last)],
                                  "path": [int(p) for p in getattr(model, "path%d" % last)],
                                  "downs": [int(d) for d in getattr(model, "downs%d" % last)],
                                  "widths": [float(w) for w in getattr(model, "widths%d" % last)]} for last in (0, 1, 2)}
    # the decoder's side effect on the caller's tensors: which alpha entries were set to -inf
    rec["alphas_neg_inf"] = [[int(i) for i in torch.nonzero(torch.isinf(a).flatten()).flatten()] for a in alphas]
    rec["structures"] = {}
    model.train(training)
    for lasts in LASTS_CHOICES:
        alphas, betas, ratios = clone_params(case)
        try:
            m = Net(alphas, betas, ratios, num_classes=19, layers=case["layers"], Fch=12, width_mult_list=WML,
                    stem_head_width=case["stem_head_width"], ignore_skip=case["ignore_skip"])
            m.train(training)
            m.build_structure(list(lasts))
            rec["structures"][",".join(map(str, lasts))] = describe(m)
        except Exception as e:  # noqa: BLE001
            rec["structures"][",".join(map(str, lasts))] = {"error": type(e).__name__}
    return rec


def main():
    ns = ref_harness.load_reference("train", "model_seg", "operations", "seg_oprs")
    Net = ns.model_seg.Network_Multi_Path_Infer
    for mod in (ns.operations, ns.seg_oprs):     # the two reference modules that hold a `latency_lookup_table` global
        assert isinstance(mod.latency_lookup_table, dict)
        mod.latency_lookup_table = SyntheticLatencyTable()
    out = {"n_cases": 0, "cases": {}}
    for seed in range(1000, 1120):
        case = draw_case(seed)
        out["cases"][str(seed)] = {"training": bool(seed % 2), "rec": run_case(Net, case, bool(seed % 2))}
    out["n_cases"] = len(out["cases"])
    n_err = sum(1 for c in out["cases"].values() if "ctor_error" in c["rec"])
    n_serr = sum(1 for c in out["cases"].values() for s in c["rec"].get("structures", {}).values() if "error" in s)
    print("cases %d, constructor rejections %d, structure rejections %d" % (out["n_cases"], n_err, n_serr))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "decode_fuzz.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
