set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_latest.json; cat gpurun_out/bench_latest.json; tail -3 gpurun_out/bench.err
bash tools/profile_r1.sh
python tools/build_times.py > gpurun_out/build_times.txt 2>&1
python tests/config_table.py > gpurun_out/config_table.txt 2>&1; tail -8 gpurun_out/config_table.txt
