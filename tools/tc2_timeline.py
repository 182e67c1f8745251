#!/usr/bin/env python
"""Phase timeline (ns, %globaltimer) of the row-strip conv kernel for one layer: first and last CTA."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasterseg_b200 import _lib  # noqa: E402
from fasterseg_b200 import functional as F_  # noqa: E402

ci, co, h, w = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (128, 128, 128, 256))]
dev = torch.device("cuda")
x = F_.empty_nhwc(1, ci, h, w, dev).normal_()
y = F_.empty_nhwc(1, co, h, w, dev)
wp = F_.pack_conv_weight(torch.randn(co, ci, 3, 3, device=dev) * 0.05, ci, co, 3)
sc, sh = torch.rand(co, device=dev) + 0.5, torch.randn(co, device=dev) * 0.1
for _ in range(3):
    F_.conv_fwd(x, wp, co, 3, 1, 1, sc, sh, relu=True, out=y)
torch.cuda.synchronize()
buf = torch.zeros(128, dtype=torch.int64, device=dev)
_lib.lib().fsb_debug_set_buffer(C.c_void_p(buf.data_ptr()))
F_.conv_fwd(x, wp, co, 3, 1, 1, sc, sh, relu=True, out=y)
torch.cuda.synchronize()
_lib.lib().fsb_debug_set_buffer(None)
b = buf.cpu().tolist()
names = {0: "entry", 1: "prologue done", 2: "A[kc0] landed", 12: "A[kc1] landed", 40: "all MMAs issued", 41: "accumulators complete",
         42: "epilogue done"}
for base, label in ((0, "first CTA"), (64, "last CTA")):
    t0 = b[base]
    print("== %s" % label)
    for i in range(43):
        if b[base + i]:
            nm = names.get(i, "B tile kc%d tap%d ready" % ((i - 3) // 10, (i - 3) % 10))
            print("  %2d %-26s +%7d ns" % (i, nm, b[base + i] - t0))
