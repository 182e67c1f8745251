import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import fasterseg_oracle as orc
from tests import helpers as H
from tests.test_boundary_cpu import _build_student
from tests.test_student_gpu import _load_seeded
from fasterseg_b200 import functional as F_
model, g = _build_student(1, training=True)
model = model.cuda().train()
sd = _load_seeded(model, g, 2025, key="state_dict_shapes_train")
st, _ = H.student_structure(1)
x = orc.random_input((2, 3, 192, 384), seed=100)
P = orc.Params({k: v.clone() for k, v in sd.items()})
outs = {}
def hook(name):
    def f(m, i, o):
        outs[name] = F_.to_nchw(o.detach(), torch.float32).cpu() if F_.is_nhwc_half(o) else o.detach().float().cpu()
    return f
for i in range(3): model.stem[i].register_forward_hook(hook("stem.%d" % i))
for k, c in model.cells.items(): c.register_forward_hook(hook("cells." + k))
with torch.no_grad():
    model(x.cuda())
    y = orc.conv_norm(x, P.sub("stem.0"), 3, 2, 1, True); print("stem.0", H.rel_err(outs["stem.0"].numpy(), y.numpy()))
    y = orc.basic_residual_2x(y, P.sub("stem.1"), 2, True); print("stem.1", H.rel_err(outs["stem.1"].numpy(), y.numpy()))
    y = orc.basic_residual_2x(y, P.sub("stem.2"), 2, True); print("stem.2", H.rel_err(outs["stem.2"].numpy(), y.numpy()))
    outputs = [y] * 2
    for layer, groups in enumerate(st.branch_groups):
        for grp in groups:
            spec = st.cells["%d-%d" % (layer, grp[0])]
            o = orc.OP_FUNCS[spec.op](outputs[grp[0]], P.sub("cells.%d-%d._op._op" % (layer, grp[0])), 2 if spec.down else 1, True)
            for b in grp: outputs[b] = o
            got = outs["cells.%d-%d" % (layer, grp[0])]
            print("cell %d-%d op%d %s" % (layer, grp[0], spec.op, tuple(o.shape)), H.rel_err(got.numpy(), o.numpy()))
