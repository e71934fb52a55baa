#!/bin/bash
# tools/final_r2.sh -- the round's closing evidence in one gpurun call: GPU test suite, bench line, config table,
# per-depth traversal log, then tools/profile_r2.sh (ncu launch list, traffic, --set full captures).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu.txt
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python tests/config_table.py > gpurun_out/r2_config_table.txt 2>&1
KFRT_TRACE_LOG=1 KFRT_INSTANCE_SUBTREES=0 python tools/counters.py million 1920 1080 16 2>&1 | grep "\[kfrt\]" | tail -18 > gpurun_out/r2_trace_by_depth.txt
python tools/build_times.py > gpurun_out/r2_build_times.txt 2>&1
tools/profile_r2.sh > gpurun_out/r2_profile.log 2>&1
cat gpurun_out/r2_pytest_gpu.txt; tail -c 600 gpurun_out/r2_bench.json; tail -8 gpurun_out/r2_config_table.txt
