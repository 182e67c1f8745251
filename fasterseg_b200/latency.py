"""Wall-clock latency of a module on the current CUDA device -- the reference's measurement protocol
(tools/utils/darts_utils.py:182-223: 10 warm-ups, doubling loop until >= 1 s, then int(FPS*6) timed iterations),
timed with CUDA events on the launching stream instead of host clocks around two synchronizes."""
import torch


def compute_latency_ms(model, input_size, iterations=None, device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("compute_latency_ms needs a CUDA device (there is no CPU fallback)")
    model.eval()
    model = model.cuda()
    x = torch.randn(*input_size).cuda()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(n):
        torch.cuda.synchronize()
        start.record()
        for _ in range(n):
            model(x)
        stop.record()
        stop.synchronize()
        return start.elapsed_time(stop) / 1000.0

    with torch.no_grad():
        for _ in range(10):
            model(x)
        if iterations is None:
            elapsed, iterations = 0.0, 100
            while elapsed < 1:
                elapsed = timed(iterations)
                iterations *= 2
            fps = iterations / elapsed
            iterations = int(fps * 6)
        elapsed = timed(iterations)
    return elapsed / iterations * 1000
