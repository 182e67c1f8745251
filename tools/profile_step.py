import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["search_step_bench.py", "--mode", "pretrain", "--steps", "1", "--warmup", "1", "--layers", "16"]
import tools.search_step_bench as ssb
pr = cProfile.Profile()
pr.enable()
ssb.main()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
