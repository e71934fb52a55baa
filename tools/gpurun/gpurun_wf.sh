set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py $BENCHARGS 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_latest.json; cat gpurun_out/bench_latest.json; tail -5 gpurun_out/bench.err
if [ -n "$PROF" ]; then bash tools/profile_r1.sh; fi
