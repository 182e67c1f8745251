#!/usr/bin/env bash
# 2-GPU session: data-parallel captured passes after serialising the peer exchanges on one ordered stream.
set -u
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29541 tools/dp_graph_check.py > gpurun_out/r2s10_dp_graph_check.log 2>&1; echo "dp_graph_check rc=$?"
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2s10_dp_graph_check.log | tail -8 | cut -c1-900
timeout 300 $RUN --master-port 29542 tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s10_pretrain_2gpu.log 2>&1; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2s10_pretrain_2gpu.log | tail -3 | cut -c1-700
timeout 300 $RUN --master-port 29543 tools/search_step_bench.py --mode search --steps 5 --warmup 2 > gpurun_out/r2s10_search_2gpu.log 2>&1; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2s10_search_2gpu.log | tail -3 | cut -c1-700
timeout 600 $RUN --master-port 29544 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2s10_bench_2gpu.json 2> gpurun_out/r2s10_bench_2gpu.err; echo "bench2 rc=$?"; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2s10_bench_2gpu.err | tail -3; cut -c1-300 gpurun_out/r2s10_bench_2gpu.json
