set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py --steps 2 --warmup 1 --spp 16 --no-cpu-baseline 2>&1 | tail -2
