#!/usr/bin/env bash
# GPU session 16: flat step tail (clip + SGD kernels): parity tests, pretrain / search step with and without it.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optim_gpu.py -q 2>&1 | tail -15 | cut -c1-250
timeout 300 python tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s16_pretrain_flat.log 2>&1; tail -1 gpurun_out/r2s16_pretrain_flat.log | cut -c1-500
timeout 300 python tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 --flat-optim 0 > gpurun_out/r2s16_pretrain_torch.log 2>&1; tail -1 gpurun_out/r2s16_pretrain_torch.log | cut -c1-300
timeout 300 python tools/search_step_bench.py --mode search --steps 5 --warmup 2 > gpurun_out/r2s16_search_flat.log 2>&1; tail -1 gpurun_out/r2s16_search_flat.log | cut -c1-300
timeout 300 python tools/step_census.py pretrain 2>&1 | grep "^step" | head -3
