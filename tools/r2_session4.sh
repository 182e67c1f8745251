#!/usr/bin/env bash
# GPU session 4: multi-stream captured passes + BASELINE-size parity tests.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_supernet_gpu.py tests/test_baseline_sizes_gpu.py -x -q -s > gpurun_out/r2s4_pytest.log 2>&1
grep -n "step 0\|step 1\|C2 \|c3:\|c5:\|passed\|failed\|Error" gpurun_out/r2s4_pytest.log | cut -c1-400 | tail -40
timeout 300 python tools/search_step_bench.py --mode pretrain --steps 10 --warmup 3 > gpurun_out/r2s4_pretrain_graph_streams.log 2>&1; tail -1 gpurun_out/r2s4_pretrain_graph_streams.log
FSB_GRAPH_STREAMS=0 timeout 300 python tools/search_step_bench.py --mode pretrain --steps 10 --warmup 3 > gpurun_out/r2s4_pretrain_graph_nostreams.log 2>&1; tail -1 gpurun_out/r2s4_pretrain_graph_nostreams.log
timeout 300 python tools/search_step_bench.py --mode search --steps 6 --warmup 2 > gpurun_out/r2s4_search_graph_streams.log 2>&1; tail -1 gpurun_out/r2s4_search_graph_streams.log
