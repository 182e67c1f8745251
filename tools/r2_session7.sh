#!/usr/bin/env bash
# GPU session 7: MMA issue-rate sweeps (accumulator round-robin / commit cost / ring depth), first run of the CTA-pair conv kernel
# (parity tests forced on, then the per-layer table with it), in-situ frame timeline.
set -u
mkdir -p gpurun_out
timeout 200 ./tools/umma_rate > gpurun_out/r2s7_umma_rate.log 2>&1; echo "umma_rate rc=$?"; cat gpurun_out/r2s7_umma_rate.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k tc4 > gpurun_out/r2s7_pytest_tc4.log 2>&1; echo "pytest tc4 rc=$?"; grep -v "^$" gpurun_out/r2s7_pytest_tc4.log | tail -40 | cut -c1-300
FSB_CONV_TC4=1 timeout 200 python tools/conv_bench.py > gpurun_out/r2s7_conv_bench_tc4.log 2>&1; echo "conv_bench rc=$?"; tail -30 gpurun_out/r2s7_conv_bench_tc4.log
timeout 200 python tools/frame_profile.py > gpurun_out/r2s7_frame_profile.log 2>&1; tail -100 gpurun_out/r2s7_frame_profile.log
