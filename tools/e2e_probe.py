#!/usr/bin/env python
"""Where does the end-to-end frame time go?  H2D / D2H bandwidth from pinned memory, pipeline throughput vs depth."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import H, W, synth_weights_  # noqa: E402
from fasterseg_b200 import zoo  # noqa: E402
from fasterseg_b200.runtime import InferencePipeline  # noqa: E402

dev = torch.device("cuda")
for dt, name in ((torch.float32, "fp32"), (torch.float16, "fp16"), (torch.uint8, "uint8")):
    h = torch.empty((1, 3, H, W), dtype=dt).pin_memory()
    d = torch.empty((1, 3, H, W), dtype=dt, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    en.record()
    en.synchronize()
    ms = st.elapsed_time(en) / 20
    print("H2D %s frame: %.3f ms  (%.1f GB/s)" % (name, ms, h.numel() * h.element_size() / ms / 1e6))
model = zoo.build_network(1)
synth_weights_(model)
model = model.to(dev).eval()
x = torch.randn(1, 3, H, W, device=dev)
frames = [torch.randn(1, 3, H, W).pin_memory() for _ in range(4)]
for depth in (1, 2, 3, 4):
    pipe = InferencePipeline(model, x, mode="labels", depth=depth)
    pipe.run(frames[i % 4] for i in range(20))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    pipe.run(frames[i % 4] for i in range(n))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("pipeline depth %d: %.1f FPS (%.3f ms/frame)" % (depth, n / dt, dt / n * 1e3))
    del pipe
