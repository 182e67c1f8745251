#!/usr/bin/env bash
# 2-GPU session: push/pull peer exchange with 8 pull streams.
set -u
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
MODE=pretrain timeout 300 $RUN --master-port 29551 tools/dp_graph_check.py > gpurun_out/r2s11_dp_graph_check_pretrain.log 2>&1; echo "dp_graph_check pretrain rc=$?"
grep "^{\|DP GRAPH" gpurun_out/r2s11_dp_graph_check_pretrain.log | cut -c1-900
MODE=search timeout 300 $RUN --master-port 29552 tools/dp_graph_check.py > gpurun_out/r2s11_dp_graph_check_search.log 2>&1; echo "dp_graph_check search rc=$?"
grep "^{\|DP GRAPH" gpurun_out/r2s11_dp_graph_check_search.log | cut -c1-900
timeout 300 $RUN --master-port 29553 tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s11_pretrain_2gpu.log 2>&1; grep "^{\|timed out\|Error" gpurun_out/r2s11_pretrain_2gpu.log | tail -3 | cut -c1-700
timeout 300 $RUN --master-port 29554 tools/search_step_bench.py --mode search --steps 5 --warmup 2 > gpurun_out/r2s11_search_2gpu.log 2>&1; grep "^{\|timed out\|Error" gpurun_out/r2s11_search_2gpu.log | tail -3 | cut -c1-700
