import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
from torch.profiler import ProfilerActivity, profile
from tools.search_step_bench import build, weight_params
torch.manual_seed(1); np.random.seed(1)
model = build(16)
opt = torch.optim.SGD(weight_params(model), lr=0.02, momentum=0.9, weight_decay=5e-4)
x = torch.randn(3, 3, 256, 512, device="cuda"); t = torch.randint(0, 19, (3, 32, 64), device="cuda")
def step():
    opt.zero_grad(); loss = model._loss(x, t, True); loss.backward(); nn.utils.clip_grad_norm_(model.parameters(), 5); opt.step()
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.time_range.end - e.time_range.start for e in ev)
span = max(e.time_range.end for e in ev) - min(e.time_range.start for e in ev)
agg = {}
for e in ev:
    k = e.name.split("(")[0].replace("void ", "")[:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
print("GPU events %d, sum of durations %.1f ms, span %.1f ms" % (len(ev), tot / 1e3, span / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print("%-62s n=%6d total %8.1f ms" % (k, v[0], v[1] / 1e3))
