"""bench.py contract without a GPU: the reference arm prints ONE JSON line with the agreed keys (same metric / unit / workload
as our arm), ranks other than 0 stay silent, and our arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, "bench.py"] + args, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fasterseg_student_fps_1024x2048" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    # "reference" when the reference tree (or build()'s verbatim copy under oracle/_ref) is importable, the oracle port otherwise
    from oracle import ref_harness
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_harness.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "1x3x1024x2048" in d["config"]["workload"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.WORKLOAD          # same workload string as our arm


def test_reference_arm_falls_back_to_the_port_without_a_reference_tree():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"], env={"FASTERSEG_REFERENCE": "/nonexistent"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0])["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "3", "--warmup", "3"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_default_run_cpu_baseline_leg(monkeypatch):
    """`cpu_baseline` of our arm's line (bench.student_cpu_baseline): OUR model's state_dict loads into the unmodified reference network
    key for key (kind "reference"; the oracle port where no reference tree exists), thread count probed, sample bounded.  Small frame
    so that the leg runs in seconds here; the bench itself uses 1024x2048."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from fasterseg_b200 import zoo
    from oracle import ref_harness
    monkeypatch.setattr(bench, "H", 64)
    monkeypatch.setattr(bench, "W", 128)
    model = zoo.build_network(1)
    bench.synth_weights_(model)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    threads = torch.get_num_threads()
    try:
        out = bench.student_cpu_baseline(sd, seconds=0.05)
    finally:
        torch.set_num_threads(threads)
    assert out["kind"] == ("reference" if ref_harness.reference_available() else "port")
    assert out["value"] > 0 and out["cores"] >= 1 and out["unit"] == "frames/s" and "1x3x64x128" in out["sample"]


def test_numa_binding_is_best_effort_and_keeps_a_subset():
    sys.path.insert(0, ROOT)
    import bench
    before = os.sched_getaffinity(0)
    try:
        n = bench.bind_to_one_numa_node()
        after = os.sched_getaffinity(0)
        assert after <= before and len(after) >= 1
        assert n in (0, len(after))
    finally:
        for tid in os.listdir("/proc/self/task"):
            try:
                os.sched_setaffinity(int(tid), before)
            except OSError:
                pass
