#!/usr/bin/env bash
# GPU session 6 (round 2, second builder session): MMA issue-rate micro-benchmark, whole GPU suite at HEAD, default bench line.
set -u
mkdir -p gpurun_out
timeout 120 ./tools/umma_rate > gpurun_out/r2s6_umma_rate.log 2>&1; echo "umma_rate rc=$?"; cat gpurun_out/r2s6_umma_rate.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s6_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2s6_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2s6_bench.json 2> gpurun_out/r2s6_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s6_bench.err; cat gpurun_out/r2s6_bench.json
