"""Frame-level roofline bookkeeping (fasterseg_b200/roofline.py) on the CPU stand-in backend: the launch list of one student
frame at BASELINE configs[1] must reproduce SURVEY section 8(a)'s totals (45 convs incl. the stem, 55.54 GFLOP, 74 launches)."""
import torch

from fasterseg_b200 import roofline
from tests import cpu_backend
from tests.test_boundary_cpu import _build_student


def test_student_frame_launch_list_and_sigma_roofline():
    pristine = roofline.F_.conv_fwd
    with cpu_backend.installed():
        model, _ = _build_student(1)
        model.eval()
        model.logits_dtype = torch.float16
        x = torch.zeros(1, 3, 1024, 2048)
        with torch.no_grad():
            recs = roofline.trace_launches(lambda: model(x))
            recs_lab = roofline.trace_launches(lambda: model.predict_labels(x))
    kinds = [r["kernel"] for r in recs]
    assert len(recs) == 74 and kinds.count("conv") + kinds.count("stem_conv") == 45 and kinds.count("bilinear") == 25
    assert kinds[-1] == "upsample_logits" and [r["kernel"] for r in recs_lab][-1] == "upsample_argmax"
    s = roofline.sigma_roofline(recs, tensor_tflops=1694.0, hbm_gbs=6568.7)
    assert abs(s["gflop"] - 55.54) < 0.01
    assert 70.0 < s["sum_us"] < 85.0 and s["tensor_bound_launches"] == 18
    # the fused upsample+argmax moves 3.3 MB (1.2 MB logits in, 2.1 MB labels out) instead of 80.9 MB of fp16 logits
    assert recs_lab[-1]["bytes"] < recs[-1]["bytes"] / 20
    assert roofline.trace_launches(lambda: None) == []
    assert roofline.F_.conv_fwd is pristine      # instrumentation (and the stand-in backend) removed again
