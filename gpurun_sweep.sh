python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ri in 1 4 8 16 24; do echo "== refill $ri"; KFRT_REFILL_IDLE=$ri python tools/counters.py million 0 0 8 2>&1 | tail -1; done
