mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/build_times.py 2>&1 | tee gpurun_out/build_times.txt
