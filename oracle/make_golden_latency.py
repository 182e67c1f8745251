#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- goldens of the supernet's expected-latency model (`Network_Multi_Path.forward_latency`,
search/model_search.py:361-475) from the UNMODIFIED reference, over a synthetic lookup table
(oracle/make_golden_decode.py::SyntheticLatencyTable, so no reference data file is needed to reproduce them) and random
architecture parameters: every combination of the alpha / beta / ratio switches, both architectures (teacher: forced-max
widths, student: gumbel-sampled widths), several layer counts.  Written to tests/golden/supernet_latency.json.
Run in the build container:  python oracle/make_golden_latency.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness  # noqa: E402
from oracle.make_golden_decode import SyntheticLatencyTable  # noqa: E402

WML = [4. / 12, 6. / 12, 8. / 12, 10. / 12, 1.]
CASES = [{"seed": 100 + n, "layers": layers, "hw": hw} for n, (layers, hw) in
         enumerate([(5, (256, 512)), (6, (224, 448)), (6, (1024, 2048)), (8, (256, 512)), (9, (512, 1024)), (16, (1024, 2048))])]
FLAGS = [(a, b, r) for a in (True, False) for b in (True, False) for r in (True, False)]


def build(Net, layers):
    return Net(19, layers, nn.CrossEntropyLoss(ignore_index=255), Fch=12, width_mult_list=WML, prun_modes=['max', 'arch_ratio'],
               stem_head_width=[(1, 1), (8. / 12, 8. / 12)])


def randomise_arch(model, seed):
    """random architecture parameters (numpy MT19937 -> identical on every platform), in named_parameters order"""
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.split("_")[0] in ("alpha", "beta", "ratio"):
                p.copy_(torch.from_numpy(rs.standard_normal(tuple(p.shape)).astype(np.float32)))


def evaluate(model, case):
    """{"a<idx>.<alpha><beta><ratio>": latency} -- called with the reference model here and with ours in the tests"""
    out = {}
    H, W = case["hw"]
    for arch_idx in (0, 1):
        for flags in FLAGS:
            model.arch_idx, model.prun_mode = arch_idx, None
            torch.manual_seed(case["seed"] * 7 + arch_idx)     # gumbel noise of the arch_ratio sampling
            np.random.seed(case["seed"] * 11 + arch_idx)
            with torch.no_grad():
                lat = model.forward_latency((3, H, W), alpha=flags[0], beta=flags[1], ratio=flags[2])
            out["a%d.%d%d%d" % ((arch_idx,) + tuple(int(f) for f in flags))] = float(lat)
    return out


def main():
    ns = ref_harness.load_reference("search", "slimmable_ops", "operations", "seg_oprs", "genotypes", "model_search")
    for mod in (ns.operations, ns.seg_oprs):
        assert isinstance(mod.latency_lookup_table, dict)
        mod.latency_lookup_table = SyntheticLatencyTable()
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self      # model_search.py:373-384 hard-codes .cuda(); generator process only
    golden = {}
    for case in CASES:
        model = build(ns.model_search.Network_Multi_Path, case["layers"])
        randomise_arch(model, case["seed"])
        golden[str(case["seed"])] = evaluate(model, case)
        print(case, {k: round(v, 4) for k, v in list(golden[str(case["seed"])].items())[:3]})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "supernet_latency.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
