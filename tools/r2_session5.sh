#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_supernet_gpu.py -x -q -s -k captured > gpurun_out/r2s5_pytest_captured.log 2>&1
grep -n "step 0\|step 1\|passed\|failed\|Error" gpurun_out/r2s5_pytest_captured.log | cut -c1-300 | tail
timeout 900 python -m pytest tests/test_baseline_sizes_gpu.py tests/test_student_gpu.py -q -s > gpurun_out/r2s5_pytest_baseline.log 2>&1
grep -n "C2 \|c3:\|c5:\|CPU oracle\|passed\|failed\|Error\|assert" gpurun_out/r2s5_pytest_baseline.log | cut -c1-330 | tail -40
timeout 300 python tools/search_step_bench.py --mode pretrain --steps 10 --warmup 3 > gpurun_out/r2s5_pretrain.log 2>&1; tail -1 gpurun_out/r2s5_pretrain.log
timeout 300 python tools/search_step_bench.py --mode search --steps 6 --warmup 2 > gpurun_out/r2s5_search.log 2>&1; tail -1 gpurun_out/r2s5_search.log
