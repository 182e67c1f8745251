#!/usr/bin/env bash
# 8-GPU session: data-parallel captured supernet SEARCH step (architect step + weight step).
set -u
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 260 $RUN --master-port 29591 tools/search_step_bench.py --mode search --steps 5 --warmup 2 > gpurun_out/r2s19_search_8gpu.log 2>&1; echo "search8 rc=$?"
grep "^{\|timed out\|Error\|error" gpurun_out/r2s19_search_8gpu.log | tail -4 | cut -c1-900
