#!/usr/bin/env bash
# Next round, 2-GPU session: validate and measure the library-owned NCCL exchange (FSB_NATIVE_DP=1: SyncBN statistics
# all-reduced on the stream INSIDE the fused training units) against the torch.distributed path.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_dp_session.sh'
set -u
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
FSB_NATIVE_DP=1 timeout 200 $RUN --master-port 29521 tools/dp_check.py > gpurun_out/r2_dp_check_native.log 2>&1
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2_dp_check_native.log | tail -14
timeout 200 $RUN --master-port 29522 tools/search_step_bench.py --mode pretrain --steps 3 --warmup 1 > gpurun_out/r2_pretrain_2gpu_torchdist.log 2>&1
tail -1 gpurun_out/r2_pretrain_2gpu_torchdist.log
FSB_NATIVE_DP=1 timeout 200 $RUN --master-port 29523 tools/search_step_bench.py --mode pretrain --steps 3 --warmup 1 > gpurun_out/r2_pretrain_2gpu_native.log 2>&1
tail -1 gpurun_out/r2_pretrain_2gpu_native.log
