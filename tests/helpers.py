"""Shared helpers for the parity tests: rebuild golden cases (seeded inputs / weights) and
run the oracle on them.  Test infrastructure only."""
import json
import os

import numpy as np
import torch

from oracle import fasterseg_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def gen_x(seed, shape):
    return np.random.RandomState(seed * 3 + 1).standard_normal(shape).astype(np.float32)


def gen_gy(seed, shape):
    return np.random.RandomState(seed * 3 + 2).standard_normal(shape).astype(np.float32)


def case_state_dict(meta):
    shapes = {k: tuple(v) for k, v in meta["shapes"].items()}
    return orc.random_state_dict(shapes, seed=meta["seed"])


def oracle_run_op(meta, x, sd):
    """Run the oracle restatement of the op described by an ops_meta.json entry."""
    P = orc.Params(sd)
    cls = meta["cls"]
    training = meta["training"]
    ratio = None
    if meta.get("slimmable") and meta["ratio"] is not None:
        ratio = orc.Ratio(meta["ratio"][0], meta["ratio"][1], orc.WIDTH_MULT_LIST)
    stride = meta.get("stride", 1)
    if cls == "ConvNorm":
        return orc.conv_norm(x, P, meta["kernel_size"], stride, None, training, ratio)
    if cls == "Head":
        return orc.head(x, P, training)
    if cls == "FeatureFusion":
        return orc.feature_fusion(x, P, training)
    idx = ["FactorizedReduce", "BasicResidual1x", "BasicResidual_downup_1x", "BasicResidual2x",
           "BasicResidual_downup_2x"].index(cls)
    if idx == 0:
        return orc.factorized_reduce(x, P, stride, training, ratio, slimmable=meta["slimmable"])
    return orc.OP_FUNCS[idx](x, P, stride, training, ratio)


def student_structure(arch_idx, lasts=None):
    g = load_json("genotypes.json")["arch_%d" % arch_idx]
    a = g["arch"]
    t = lambda k: torch.tensor(a[k], dtype=torch.float32)
    alphas = [t("alpha_%d_%d" % (arch_idx, s)) for s in range(3)]
    betas = [None, t("beta_%d_1" % arch_idx), t("beta_%d_2" % arch_idx)]
    ratios = [t("ratio_%d_%d" % (arch_idx, s)) for s in range(3)]
    shw = (1.0, 1.0) if arch_idx == 0 else (8. / 12, 8. / 12)
    if lasts is None:
        lasts = g["lasts"]
    st = orc.StudentStructure(alphas, betas, ratios, lasts, stem_head_width=shw, ignore_skip=(arch_idx == 0))
    return st, g


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
