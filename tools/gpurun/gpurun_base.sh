set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py ${BENCHARGS:-} 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_latest.json; cat gpurun_out/bench_latest.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cat gpurun_out/bench_ref.json
