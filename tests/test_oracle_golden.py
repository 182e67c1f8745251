"""Pin the CPU oracle (oracle/fasterseg_oracle.py) to golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import fasterseg_oracle as orc
from tests import helpers as H

FP32_TOL = 2e-5  # oneDNN vs our summation order, fp32


def test_make_divisible_kat():
    for c, w, expect in H.load_json("make_divisible.json"):
        assert orc.make_divisible(c * w) == expect, (c, w)


def test_bilinear_matches_reference():
    z = H.load_npz("bilinear.npz")
    n = len([k for k in z.files if k.endswith("/x")])
    assert n >= 9
    for i in range(n):
        x, y = z["%d/x" % i], z["%d/y" % i]
        got = orc.bilinear_ac(torch.from_numpy(x), y.shape[2:]).numpy()
        np.testing.assert_allclose(got, y, rtol=0, atol=3e-6)


OPS_META = H.load_json("ops_meta.json")


@pytest.mark.parametrize("name", sorted(OPS_META))
def test_op_forward_backward_matches_reference(name):
    meta = OPS_META[name]
    z = H.load_npz("ops.npz")
    x = torch.from_numpy(H.gen_x(meta["seed"], tuple(meta["x_shape"]))).requires_grad_(meta["training"])
    sd = H.case_state_dict(meta)
    if meta["training"]:
        for k, v in sd.items():
            if "running" not in k:
                v.requires_grad_(True)
    y = H.oracle_run_op(meta, x, sd)
    ref = z[name + "/y"]
    assert tuple(y.shape) == ref.shape
    assert H.rel_err(y.detach().numpy(), ref) < FP32_TOL
    if not meta["training"] or not meta["shapes"]:  # identity skip has no params / no golden grads
        return
    gy = torch.from_numpy(H.gen_gy(meta["seed"], ref.shape))
    y.backward(gy)
    assert H.rel_err(x.grad.numpy(), z[name + "/gx"]) < 1e-4
    n_grads = 0
    for k in z.files:
        if k.startswith(name + "/grad:"):
            key = k.split("grad:")[1]
            assert sd[key].grad is not None, key
            assert H.rel_err(sd[key].grad.numpy(), z[k]) < 2e-4, key
            n_grads += 1
        if k.startswith(name + "/after:"):
            key = k.split("after:")[1]
            np.testing.assert_allclose(sd[key].detach().numpy(), z[k], rtol=1e-5, atol=1e-6)
    if meta["shapes"]:
        assert n_grads > 0
    # parameters the reference leaves without grad (dead USBatchNorm2d weight/bias, unused widths)
    ref_keys = {k.split("grad:")[1] for k in z.files if k.startswith(name + "/grad:")}
    for k, v in sd.items():
        if v.grad is not None:
            assert k in ref_keys, "oracle used a parameter the reference does not: " + k


def test_genotype_decode_kat():
    g = H.load_json("genotypes.json")
    for arch_idx in (0, 1):
        st, ref = H.student_structure(arch_idx)
        d = st.describe()
        assert d["lasts"] == ref["lasts"]
        for bi, last in enumerate(st.lasts):
            r = ref["decoded"][str(last)]
            assert d["ops"][bi] == r["ops"]
            assert d["paths"][bi] == r["path"]
            assert d["downs"][bi] == r["downs"]
            assert np.allclose(d["widths"][bi], r["widths"])
        assert d["branch_groups"] == ref["branch_groups"]
        assert (d["ch_16"], d["ch_8_2"], d["ch_8_1"]) == (ref["ch_16"], ref["ch_8_2"], ref["ch_8_1"])
        names = ["FactorizedReduce", "BasicResidual1x", "BasicResidual_downup_1x", "BasicResidual2x",
                 "BasicResidual_downup_2x"]
        for k, (ci, co, down, cls) in ref["cells"].items():
            spec = st.cells[k]
            assert (spec.c_in, spec.c_out, spec.down, names[spec.op]) == (ci, co, down, cls), k
        # state_dict key/shape set the forward consumes is a subset of the reference's, identical shapes
        shapes = orc.student_param_shapes(st, training=False)
        for k, shp in shapes.items():
            assert list(shp) == ref["state_dict_shapes"][k], k
        missing = [k for k in ref["state_dict_shapes"] if k not in shapes and "num_batches_tracked" not in k
                   and "channel_attention" not in k]
        assert missing == [], missing
    # student as in SURVEY section 8a
    st, _ = H.student_structure(1)
    assert st.lasts == [2, 1] and (st.ch_16, st.ch_8_2, st.ch_8_1) == (64, 32, 32)


def test_student_train_build_shapes():
    st, ref = H.student_structure(1)
    shapes = orc.student_param_shapes(st, training=True)
    for k, shp in shapes.items():
        assert list(shp) == ref["state_dict_shapes_train"][k], k


@pytest.mark.parametrize("arch_idx,hw", [(1, (64, 128)), (0, (64, 128)), (1, (96, 160))])
def test_student_eval_matches_reference(arch_idx, hw):
    z = H.load_npz("student.npz")
    st, g = H.student_structure(arch_idx)
    full = {k: tuple(v) for k, v in g["state_dict_shapes"].items() if not k.endswith("num_batches_tracked")}
    sd = orc.random_state_dict(full, seed=2024 + arch_idx)
    x = orc.random_input((1, 3) + hw, seed=99 + arch_idx)
    with torch.no_grad():
        y = orc.student_forward(x, sd, st, training=False).numpy()
    tag = "arch%d.%dx%d.eval" % (arch_idx, hw[0], hw[1])
    assert H.rel_err(y[:, :, ::4, ::4], z[tag + "/logits.s4"]) < 5e-5
    if tag + "/logits" in z.files:
        assert H.rel_err(y, z[tag + "/logits"]) < 5e-5
    agree = (y.argmax(1).astype(np.uint8) == z[tag + "/argmax"]).mean()
    assert agree > 0.9995, agree


@pytest.mark.parametrize("hw", [(64, 128), (192, 384)])
def test_student_train_matches_reference(hw):
    z = H.load_npz("student.npz")
    st, g = H.student_structure(1)
    full = {k: tuple(v) for k, v in g["state_dict_shapes_train"].items() if not k.endswith("num_batches_tracked")}
    sd = orc.random_state_dict(full, seed=2025)
    x = orc.random_input((2, 3) + hw, seed=100)
    with torch.no_grad():
        p8, p16, p32 = orc.student_forward(x, sd, st, training=True)
    tag = "arch1.%dx%d.train" % hw
    for name, o in (("pred8", p8), ("pred16", p16), ("pred32", p32)):
        assert H.rel_err(o.numpy()[:, :, ::4, ::4], z[tag + "/" + name + ".s4"]) < 1e-4, name
    for k in ("stem.0.conv.1.running_mean", "stem.0.conv.1.running_var", "heads8.conv_3x3.bn.running_var"):
        np.testing.assert_allclose(sd[k].numpy(), z[tag + "/after:" + k], rtol=2e-5, atol=1e-6)


def test_convnorm_gate_matches_reference():
    """BASELINE.json configs[0]."""
    z = H.load_npz("convnorm_gate.npz")
    x = orc.random_input((1, 3, 256, 512), seed=12345)
    for co in (32, 48):
        shapes = {"conv.0.weight": (co, 3, 3, 3), "conv.1.weight": (co,), "conv.1.bias": (co,),
                  "conv.1.running_mean": (co,), "conv.1.running_var": (co,)}
        for training in (False, True):
            sd = orc.random_state_dict(shapes, seed=12345 + co)
            with torch.no_grad():
                y = orc.conv_norm(x, orc.Params(sd), 3, 2, 1, training).numpy()
            tag = "co%d.%s" % (co, "train" if training else "eval")
            np.testing.assert_allclose(y[:, :, ::8, ::8], z[tag + "/sample"], rtol=1e-4, atol=2e-5)
            m = z[tag + "/moments"]
            assert abs(y.mean() - m[0]) < 1e-5 and abs(y.std() - m[1]) < 1e-5
            np.testing.assert_allclose(y.sum(axis=(0, 1, 3)), z[tag + "/rowsum"], rtol=1e-4, atol=1e-2)
