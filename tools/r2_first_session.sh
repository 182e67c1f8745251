#!/usr/bin/env bash
# First GPU session of the next round: validate and measure the persistent conv kernel (FSB_CONV_PERSIST=1) against the
# per-tile kernel.  Everything lands in gpurun_out/.  Usage (from the repo root, on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_first_session.sh'
set -u
mkdir -p gpurun_out
# 1. parity of the whole GPU suite with the persistent kernel enabled
FSB_CONV_PERSIST=1 timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_persist.log 2>&1
tail -3 gpurun_out/r2_pytest_persist.log
# 2. per-layer timings, per-tile vs persistent (the table prints the student's conv shapes)
timeout 120 python tools/conv_bench.py > gpurun_out/r2_conv_bench_pertile.log 2>&1
FSB_CONV_PERSIST=1 timeout 120 python tools/conv_bench.py > gpurun_out/r2_conv_bench_persist.log 2>&1
FSB_CONV_PERSIST=1 FSB_PERSIST_OCC=2 timeout 120 python tools/conv_bench.py > gpurun_out/r2_conv_bench_persist_occ2.log 2>&1
tail -30 gpurun_out/r2_conv_bench_persist.log; tail -30 gpurun_out/r2_conv_bench_persist_occ2.log
# 2b. barrier-free logits upsample (FSB_UPSAMPLE_V2=1): parity of everything that upsamples logits, then its time in the frame
FSB_UPSAMPLE_V2=1 timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_student_gpu.py -x -q -k "upsample or logits or eval" > gpurun_out/r2_pytest_upsample_v2.log 2>&1
tail -2 gpurun_out/r2_pytest_upsample_v2.log
FSB_UPSAMPLE_V2=1 timeout 150 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2_bench_upsample_v2.json 2> gpurun_out/r2_bench_upsample_v2.err
cut -c1-200 gpurun_out/r2_bench_upsample_v2.json; echo
# 3. headline metric both ways (no CPU baseline / supernet step: kernel comparison only)
timeout 150 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2_bench_pertile.json 2> gpurun_out/r2_bench_pertile.err
FSB_CONV_PERSIST=1 timeout 150 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2_bench_persist.json 2> gpurun_out/r2_bench_persist.err
cut -c1-400 gpurun_out/r2_bench_pertile.json; echo; cut -c1-400 gpurun_out/r2_bench_persist.json
# 4. N3: per-operator latency table of this GPU (667 entries, fast protocol)
timeout 300 python tools/build_latency_table.py --out gpurun_out/latency_lookup_table_b200.npy > gpurun_out/r2_latency_table.log 2>&1
tail -2 gpurun_out/r2_latency_table.log
# 5. tape mode (FSB_TAPE=1: one autograd node per forward pass): training parity, then the step time both ways
FSB_TAPE=1 timeout 300 python -m pytest tests/test_supernet_gpu.py tests/test_student_gpu.py tests/test_ops_gpu.py -x -q > gpurun_out/r2_pytest_tape.log 2>&1
tail -2 gpurun_out/r2_pytest_tape.log
timeout 120 python tools/search_step_bench.py --mode pretrain --steps 5 --warmup 2 > gpurun_out/r2_pretrain_step.log 2>&1; tail -1 gpurun_out/r2_pretrain_step.log
FSB_TAPE=1 timeout 120 python tools/search_step_bench.py --mode pretrain --steps 5 --warmup 2 > gpurun_out/r2_pretrain_step_tape.log 2>&1; tail -1 gpurun_out/r2_pretrain_step_tape.log
FSB_TAPE=1 timeout 120 python tools/search_step_bench.py --mode search --steps 3 --warmup 1 > gpurun_out/r2_search_step_tape.log 2>&1; tail -1 gpurun_out/r2_search_step_tape.log
