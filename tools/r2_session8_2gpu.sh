#!/usr/bin/env bash
# 2-GPU session: data-parallel captured passes (SyncBN over NVLink peer memory + one flat gradient all-reduce).
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_session8_2gpu.sh'
set -u
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29531 tools/dp_graph_check.py > gpurun_out/r2s8_dp_graph_check.log 2>&1; echo "dp_graph_check rc=$?"
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2s8_dp_graph_check.log | tail -12 | cut -c1-900
timeout 300 $RUN --master-port 29532 tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s8_pretrain_2gpu.log 2>&1; tail -1 gpurun_out/r2s8_pretrain_2gpu.log | cut -c1-700
timeout 300 $RUN --master-port 29533 tools/search_step_bench.py --mode search --steps 5 --warmup 2 > gpurun_out/r2s8_search_2gpu.log 2>&1; tail -1 gpurun_out/r2s8_search_2gpu.log | cut -c1-700
timeout 300 python tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s8_pretrain_1gpu.log 2>&1; tail -1 gpurun_out/r2s8_pretrain_1gpu.log | cut -c1-700
timeout 600 $RUN --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2s8_bench_2gpu.json 2> gpurun_out/r2s8_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2s8_bench_2gpu.err; cut -c1-600 gpurun_out/r2s8_bench_2gpu.json
