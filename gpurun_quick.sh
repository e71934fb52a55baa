python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/counters.py million 0 0 8 2>&1 | tail -4
