#!/usr/bin/env bash
# GPU session 13: split-K over a 3-CTA cluster in conv_tc (default on): kernel / operator parity, per-layer table, frame and supernet step.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_ops_gpu.py tests/test_student_gpu.py -q > gpurun_out/r2s13_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2s13_pytest.log | cut -c1-250
timeout 200 python tools/conv_bench.py > gpurun_out/r2s13_conv_bench_ksplit.log 2>&1; tail -28 gpurun_out/r2s13_conv_bench_ksplit.log
FSB_CONV_KSPLIT=0 timeout 200 python tools/conv_bench.py > gpurun_out/r2s13_conv_bench_noksplit.log 2>&1; tail -1 gpurun_out/r2s13_conv_bench_noksplit.log
timeout 300 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2s13_bench.json 2> gpurun_out/r2s13_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2s13_bench.json
timeout 300 python tools/search_step_bench.py --mode pretrain --steps 8 --warmup 3 > gpurun_out/r2s13_pretrain.log 2>&1; tail -1 gpurun_out/r2s13_pretrain.log | cut -c1-400
