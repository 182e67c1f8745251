#!/usr/bin/env bash
# GPU session 14: N-split lower bound of conv_tc (an MMA costs the same whatever its N): frame, per-layer table and pretrain step for 32/64/128.
set -u
mkdir -p gpurun_out
for nt in 32 64 128; do
  echo "== FSB_CONV_NTILE_MIN=$nt"
  FSB_CONV_NTILE_MIN=$nt timeout 200 python tools/conv_bench.py > gpurun_out/r2s14_conv_bench_nt$nt.log 2>&1; tail -1 gpurun_out/r2s14_conv_bench_nt$nt.log
  FSB_CONV_NTILE_MIN=$nt timeout 300 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2s14_bench_nt$nt.json 2> gpurun_out/r2s14_bench_nt$nt.err; grep -o '"value": [0-9.]*, "unit": "frames/s", "n_gpus": 1, "steps": 200, "warmup": 20, "ms_per_step": [0-9.]*' gpurun_out/r2s14_bench_nt$nt.json | head -1
  FSB_CONV_NTILE_MIN=$nt timeout 300 python tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s14_pretrain_nt$nt.log 2>&1; tail -1 gpurun_out/r2s14_pretrain_nt$nt.log | cut -c1-120
done
timeout 300 python -m pytest tests/test_student_gpu.py -q -k "multi_stream" 2>&1 | tail -2
