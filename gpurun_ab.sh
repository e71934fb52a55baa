mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
VARIANTS="head" bash gpurun_sweep.sh 2>&1 | tail -12
