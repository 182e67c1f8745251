#!/usr/bin/env bash
# GPU session 3: captured supernet passes (graph mode) -- parity vs the eager path, then the step time both ways.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_supernet_gpu.py -x -q -s > gpurun_out/r2s3_pytest_supernet.log 2>&1
tail -15 gpurun_out/r2s3_pytest_supernet.log
timeout 300 python tools/search_step_bench.py --mode pretrain --steps 10 --warmup 3 > gpurun_out/r2s3_pretrain_graph.log 2>&1; tail -2 gpurun_out/r2s3_pretrain_graph.log
timeout 300 python tools/search_step_bench.py --mode search --steps 6 --warmup 2 > gpurun_out/r2s3_search_graph.log 2>&1; tail -2 gpurun_out/r2s3_search_graph.log
timeout 200 python tools/search_step_bench.py --mode pretrain --steps 3 --warmup 1 --graph 0 > gpurun_out/r2s3_pretrain_eager.log 2>&1; tail -1 gpurun_out/r2s3_pretrain_eager.log
