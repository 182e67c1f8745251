#!/usr/bin/env python
"""Phase timeline (ns, %globaltimer) of CTA 0 of the tap-concatenated conv kernel (conv_tc5.cu) for one layer."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fasterseg_b200 import _lib  # noqa: E402
from fasterseg_b200 import functional as F_  # noqa: E402

ci, co, h, w = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 64, 256, 512))]
_lib.set_option("FSB_CONV_TC5", 2)
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
t0 = b[0]
print("entry 0, setup done +%d ns" % (b[1] - t0))
names = ["mma: accumulator free", "mma: first window landed", "mma: tile issued", "epi: accumulator complete", "epi: chunks done",
         "epi: staging synced"]
for lt in range(8):
    row = [b[2 + lt * 8 + i] for i in range(6)]
    if not any(row):
        break
    print("tile %d: " % lt + ", ".join("%s +%d" % (names[i].split(": ")[1] if False else names[i], row[i] - t0) for i in range(6) if row[i]))
