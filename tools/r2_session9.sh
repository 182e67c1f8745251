#!/usr/bin/env bash
# GPU session 9: division-free issue rings in every tcgen05 kernel + first run of the tap-concatenated kernel (conv_tc5).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s9_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/r2s9_pytest_gpu.log | tail -25 | cut -c1-250
timeout 200 python tools/conv_bench.py > gpurun_out/r2s9_conv_bench_default.log 2>&1; tail -28 gpurun_out/r2s9_conv_bench_default.log
FSB_CONV_TC5=1 timeout 200 python tools/conv_bench.py > gpurun_out/r2s9_conv_bench_tc5.log 2>&1; echo "conv_bench tc5 rc=$?"; tail -28 gpurun_out/r2s9_conv_bench_tc5.log
timeout 300 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2s9_bench.json 2> gpurun_out/r2s9_bench.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/r2s9_bench.json
FSB_CONV_TC5=1 timeout 300 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2s9_bench_tc5.json 2> gpurun_out/r2s9_bench_tc5.err; echo "bench tc5 rc=$?"; cut -c1-400 gpurun_out/r2s9_bench_tc5.json
