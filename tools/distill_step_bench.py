#!/usr/bin/env python
"""BASELINE.json configs[3]: teacher -> student KL-distillation train step (train/train.py:219-271), synthetic data.
  teacher (arch_0, eval, no_grad) forward; student (arch_1, train) forward -> 3 full-resolution logits;
  loss = CE(l8) + 0.2 CE(l16) + 0.2 CE(l32) + KLDiv(log_softmax(student l8), softmax(teacher l8)); backward; SGD step.
(The reference uses ProbOhemCrossEntropy2d for the CE terms -- a caller-side loss, "next" row N1; plain CE here.)
Prints one JSON line; --batch / --hw scale the per-GPU shard (default 12 x 3 x 512 x 1024)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_weights_  # noqa: E402
from fasterseg_b200 import zoo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--hw", type=int, nargs=2, default=[512, 1024])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    B, (H, W) = args.batch, args.hw
    teacher = zoo.build_network(0).cuda().eval()
    synth_weights_(teacher, 1)
    teacher.logits_dtype = torch.float16
    student = zoo.build_network(1, training=True).cuda().train()
    synth_weights_(student, 2)
    opt = torch.optim.SGD(student.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    ce = nn.CrossEntropyLoss(ignore_index=255)
    kl = nn.KLDivLoss(reduction="mean")
    x = torch.randn(B, 3, H, W, device="cuda")
    t = torch.randint(0, 19, (B, H, W), device="cuda")
    t[torch.rand(t.shape, device="cuda") < 0.05] = 255

    def step():
        opt.zero_grad()
        with torch.no_grad():
            tl = teacher(x).float()
        l8, l16, l32 = student(x)
        loss = ce(l8, t) + 0.2 * ce(l16, t) + 0.2 * ce(l32, t) + kl(F.softmax(l8, dim=1).log(), F.softmax(tl, dim=1))
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        loss = step()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    times.sort()
    dt = times[len(times) // 2]
    print(json.dumps({"metric": "distill_step_ms", "value": round(dt * 1e3, 1), "min_ms": round(times[0] * 1e3, 1), "unit": "ms/step",
                      "batch": [B, 3, H, W], "images_per_s": round(B / dt, 1), "loss": float(loss.detach()),
                      "mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))


if __name__ == "__main__":
    main()
