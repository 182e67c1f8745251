"""End-to-end parity of the derived network (BASELINE.json configs[1] family) on the GPU."""
import numpy as np
import pytest
import torch

from oracle import fasterseg_oracle as orc
from tests import helpers as H
from tests.test_boundary_cpu import _build_student

pytestmark = pytest.mark.gpu


def _load_seeded(model, g, seed, key="state_dict_shapes"):
    full = {k: tuple(v) for k, v in g[key].items() if not k.endswith("num_batches_tracked")}
    sd = orc.random_state_dict(full, seed=seed)
    own = model.state_dict()
    seen = set()
    for k in sorted(sd):  # shared cells: first key wins (same rule as oracle/make_golden.py)
        if own[k].data_ptr() in seen:
            continue
        seen.add(own[k].data_ptr())
        own[k].copy_(sd[k])
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps, m.momentum = orc.BN_EPS, orc.BN_MOMENTUM
    return sd


def _report(name, y, ref):
    y = y.astype(np.float64)
    ref = ref.astype(np.float64)
    nerr = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    maxerr = np.abs(y - ref).max() / (np.abs(ref).max() + 1e-12)
    print("%s: norm-rel %.3e  max-abs/max %.3e" % (name, nerr, maxerr))
    return nerr, maxerr


@pytest.mark.parametrize("arch_idx,hw", [(1, (64, 128)), (0, (64, 128)), (1, (96, 160))])
def test_student_eval_vs_reference_golden(arch_idx, hw):
    z = H.load_npz("student.npz")
    model, g = _build_student(arch_idx)
    model = model.cuda().eval()
    _load_seeded(model, g, 2024 + arch_idx)
    x = orc.random_input((1, 3) + hw, seed=99 + arch_idx).cuda()
    with torch.no_grad():
        y = model(x)
        lab = model.predict_labels(x)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and tuple(y.shape) == (1, 19) + hw and y.is_contiguous()
    tag = "arch%d.%dx%d.eval" % (arch_idx, hw[0], hw[1])
    yn = y.cpu().numpy()
    nerr, maxerr = _report(tag, yn[:, :, ::4, ::4], z[tag + "/logits.s4"])
    # north_star tolerance: logits within 1e-3 relative (fp16 storage, fp32 accumulate) of the reference
    assert maxerr < 1e-3 * 3 and nerr < 3e-3
    ref_lab = z[tag + "/argmax"]
    agree = (lab.cpu().numpy() == ref_lab).mean()
    print(tag, "argmax agreement with the fp32 reference: %.5f" % agree)
    assert agree > 0.995
    # bit-exact: fused upsample+argmax == argmax of our own logits
    assert np.array_equal(lab.cpu().numpy(), yn.argmax(1).astype(np.uint8))


def test_student_eval_vs_oracle_256x512():
    model, g = _build_student(1)
    model = model.cuda().eval()
    sd = _load_seeded(model, g, 7)
    st, _ = H.student_structure(1)
    x = orc.random_input((1, 3, 256, 512), seed=8)
    with torch.no_grad():
        ref = orc.student_forward(x, sd, st, training=False).numpy()
        y = model(x.cuda()).cpu().numpy()
        model.logits_dtype = torch.float16
        y16 = model(x.cuda()).float().cpu().numpy()
    nerr, maxerr = _report("student 256x512 vs oracle", y, ref)
    assert maxerr < 3e-3 and nerr < 3e-3
    nerr, maxerr = _report("student 256x512 fp16 logits vs oracle", y16, ref)
    assert maxerr < 4e-3
    agree = (y.argmax(1) == ref.argmax(1)).mean()
    print("argmax agreement %.5f" % agree)
    assert agree > 0.995
