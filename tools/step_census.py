#!/usr/bin/env python
"""Where does a captured supernet step go?  (1) host enqueue time vs device completion time per step, (2) kernel census of one step
(torch.profiler / CUPTI): busy span, sum of kernel durations per kernel family.   python tools/step_census.py [pretrain|search]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from fasterseg_b200 import parallel  # noqa: E402
from fasterseg_b200.losses import ProbOhemCrossEntropy2d  # noqa: E402
from tools.search_step_bench import build, weight_params  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "pretrain"
parallel.seed_all_ranks_identically(12345)
model = build(16, "ohem")
B, H, W = (3, 256, 512) if mode == "pretrain" else (2, 224, 448)
model._criterion = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=int(B * (H // 8) * (W // 8) // 16))
opt = torch.optim.SGD(weight_params(model), lr=0.02, momentum=0.9, weight_decay=5e-4)
x = torch.randn(B, 3, H, W, device="cuda")
t = torch.randint(0, 19, (B, H // 8, W // 8), device="cuda")


def step(stamps=None):
    def mark(name):
        if stamps is not None:
            stamps.append((name, time.perf_counter()))
    mark("start")
    opt.zero_grad()
    mark("zero_grad")
    loss = model._loss(x, t, True if mode == "pretrain" else "dir")
    mark("_loss enqueued")
    loss.backward()
    mark("backward enqueued")
    nn.utils.clip_grad_norm_(model.parameters(), 5)
    mark("clip enqueued")
    opt.step()
    mark("opt.step enqueued")
    torch.cuda.synchronize()
    mark("device done")


for _ in range(4):
    step()
for rep in range(3):
    st = []
    step(st)
    t0 = st[0][1]
    print("step %d: " % rep + ", ".join("%s +%.1f ms" % (n, (tt - t0) * 1e3) for n, tt in st[1:]))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.time_range.end - e.time_range.start for e in ev)
span = max(e.time_range.end for e in ev) - min(e.time_range.start for e in ev)
agg = {}
for e in ev:
    k = e.name.split("(")[0].replace("void ", "").replace("fsb::", "")[:56]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
print("GPU events %d, sum of durations %.1f ms, span %.1f ms" % (len(ev), tot / 1e3, span / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("%-58s n=%6d total %8.1f ms  avg %6.1f us" % (k, v[0], v[1] / 1e3, v[1] / v[0]))
