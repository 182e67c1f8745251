#!/usr/bin/env bash
# GPU session 12: whole suite, smoke, default bench line (all secondary metrics), ncu launch list of the bench command, ncu --set full of
# the two 9.66-GFLOP conv launches.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s12_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s12_pytest_gpu.log | cut -c1-200
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2s12_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2s12_smoke.log | cut -c1-200
timeout 1200 python bench.py > gpurun_out/r2s12_bench.json 2> gpurun_out/r2s12_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2s12_bench.err | cut -c1-200; cut -c1-600 gpurun_out/r2s12_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv --log-file gpurun_out/r2s12_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-supernet-step > gpurun_out/r2s12_ncu_bench.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/r2s12_launches_bench.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc5 -s 4 -c 2 -o gpurun_out/r2s12_prof_tc5_stem1c2 python tools/conv_bench.py --only 1 --reps 4 > gpurun_out/r2s12_ncu_tc5.log 2>&1; echo "ncu tc5 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -s 4 -c 2 -o gpurun_out/r2s12_prof_tc3_heads8 python tools/conv_bench.py --only 24 --reps 4 > gpurun_out/r2s12_ncu_tc3.log 2>&1; echo "ncu tc3 rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
