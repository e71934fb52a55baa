set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_latest.json; cat gpurun_out/bench_latest.json; tail -5 gpurun_out/bench.err
if [ -n "$PROF" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 6 -c 2 -f -o gpurun_out/prof_trace python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
fi
