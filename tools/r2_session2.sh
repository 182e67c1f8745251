#!/usr/bin/env bash
# GPU session 2 of round 2: deterministic statistics + channel-major conv kernel (conv_tc3).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2s2_pytest.log 2>&1
tail -5 gpurun_out/r2s2_pytest.log
timeout 300 python tools/determinism_probe.py --supernet --json gpurun_out/r2s2_determinism.json > gpurun_out/r2s2_determinism.log 2>&1
tail -6 gpurun_out/r2s2_determinism.log
FSB_CONV_TC3=0 timeout 120 python tools/conv_bench.py > gpurun_out/r2s2_conv_bench_tc3off.log 2>&1
timeout 120 python tools/conv_bench.py > gpurun_out/r2s2_conv_bench_tc3on.log 2>&1
FSB_CONV_TC3=2 timeout 120 python tools/conv_bench.py > gpurun_out/r2s2_conv_bench_tc3forced.log 2>&1
paste <(cut -c1-60 gpurun_out/r2s2_conv_bench_tc3off.log) <(cut -c45-60 gpurun_out/r2s2_conv_bench_tc3on.log) <(cut -c45-60 gpurun_out/r2s2_conv_bench_tc3forced.log) | tail -28
timeout 200 python bench.py --no-cpu-baseline --no-supernet-step > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err
cut -c1-300 gpurun_out/r2s2_bench.json; tail -3 gpurun_out/r2s2_bench.err
