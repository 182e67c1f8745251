"""cProfile of ONE captured supernet pretrain step (host side): where do the ~105 ms of enqueue time go?"""
import cProfile, pstats, sys, os, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
from fasterseg_b200 import parallel, optim as FO
from fasterseg_b200.losses import ProbOhemCrossEntropy2d
from tools.search_step_bench import build, weight_params
mode = sys.argv[1] if len(sys.argv) > 1 else "pretrain"
FO.install()
parallel.seed_all_ranks_identically(12345)
model = build(16, "ohem")
B, H, W = (3, 256, 512) if mode == "pretrain" else (2, 224, 448)
model._criterion = ProbOhemCrossEntropy2d(ignore_label=255, thresh=0.7, min_kept=int(B * (H // 8) * (W // 8) // 16))
opt = torch.optim.SGD(weight_params(model), lr=0.02, momentum=0.9, weight_decay=5e-4)
x = torch.randn(B, 3, H, W, device="cuda"); t = torch.randint(0, 19, (B, H // 8, W // 8), device="cuda")
def step():
    opt.zero_grad(); loss = model._loss(x, t, True if mode == "pretrain" else "dir"); loss.backward()
    nn.utils.clip_grad_norm_(model.parameters(), 5); opt.step()
for _ in range(4):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
step()
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue %.1f ms, device done +%.1f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
