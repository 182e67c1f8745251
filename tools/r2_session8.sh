#!/usr/bin/env bash
# GPU session 8: converged-warp MMA/TMA issue in every tcgen05 kernel (no per-instruction waterfall): true MMA rates, whole suite,
# per-layer table (default dispatch / CTA-pair kernel on), headline line.
set -u
mkdir -p gpurun_out
timeout 200 ./tools/umma_rate > gpurun_out/r2s8_umma_rate.log 2>&1; echo "umma_rate rc=$?"; cat gpurun_out/r2s8_umma_rate.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s8_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/r2s8_pytest_gpu.log | tail -30 | cut -c1-250
timeout 200 python tools/conv_bench.py > gpurun_out/r2s8_conv_bench_default.log 2>&1; tail -28 gpurun_out/r2s8_conv_bench_default.log
FSB_CONV_TC4=1 timeout 200 python tools/conv_bench.py > gpurun_out/r2s8_conv_bench_tc4.log 2>&1; echo "conv_bench tc4 rc=$?"; tail -28 gpurun_out/r2s8_conv_bench_tc4.log
timeout 300 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2s8_bench.json 2> gpurun_out/r2s8_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2s8_bench.json
