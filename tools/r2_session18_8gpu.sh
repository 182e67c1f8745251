#!/usr/bin/env bash
# 8-GPU session: data-parallel captured supernet pretrain step (SyncBN over NVLink peer memory, one flat gradient all-reduce, flat step tail).
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 400 -- 'bash tools/r2_session18_8gpu.sh'
set -u
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 230 $RUN --master-port 29581 tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s18_pretrain_8gpu.log 2>&1; echo "pretrain8 rc=$?"
grep "^{\|timed out\|Error\|error" gpurun_out/r2s18_pretrain_8gpu.log | tail -4 | cut -c1-900
