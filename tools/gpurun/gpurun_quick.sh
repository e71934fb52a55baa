python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/counters.py million 0 0 32 2>&1 | tail -2
