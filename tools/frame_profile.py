#!/usr/bin/env python
"""In-situ per-kernel timeline of one student frame (CUDA-graph replay) via CUPTI (torch.profiler): durations with a
warm pipeline and the gaps between kernels, which the serialised cold-cache ncu pass cannot show."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import H, W, synth_weights_  # noqa: E402
from fasterseg_b200 import zoo  # noqa: E402
from fasterseg_b200.runtime import GraphedInference  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "logits"
    dev = torch.device("cuda")
    model = zoo.build_network(1)
    synth_weights_(model)
    model = model.to(dev).eval()
    x = torch.randn(1, 3, H, W, device=dev)
    runner = GraphedInference(model, x, mode=mode, logits_dtype=torch.float16)
    for _ in range(10):
        runner.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            runner.replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
    ev.sort(key=lambda e: e.time_range.start)
    per = len(ev) // 5
    frame = ev[-per:]
    t0 = frame[0].time_range.start
    prev_end = t0
    rows = []
    for i, e in enumerate(frame):
        s, en = e.time_range.start, e.time_range.end
        rows.append((i, e.name.split("(")[0].replace("void ", "").replace("fsb::", "")[:40], en - s, s - prev_end, s - t0))
        prev_end = max(prev_end, en)
    tot = frame[-1].time_range.end - t0
    busy = sum(r[2] for r in rows)
    for r in rows:
        print("%3d %-40s dur %7.1f us  gap %6.1f us  t=%7.1f" % r)
    print("frame span %.1f us, sum of kernel durations %.1f us, kernels %d" % (tot, busy, len(rows)))
    agg = {}
    for r in rows:
        a = agg.setdefault(r[1], [0, 0.0])
        a[0] += 1
        a[1] += r[2]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-40s n=%3d total %8.1f us" % (k, v[0], v[1]))


if __name__ == "__main__":
    main()
