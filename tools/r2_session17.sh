#!/usr/bin/env bash
# GPU session 17: verification of the tree at the end of round 2 -- whole GPU test suite, smoke, both bench arms.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-300
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2s17_bench.json 2> gpurun_out/r2s17_bench.err; grep '^{' gpurun_out/r2s17_bench.json | tail -1 | cut -c1-3000
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s17_bench_reference.json 2> gpurun_out/r2s17_bench_reference.err; grep '^{' gpurun_out/r2s17_bench_reference.json | tail -1 | cut -c1-1200
tail -3 gpurun_out/r2s17_bench.err | cut -c1-300
