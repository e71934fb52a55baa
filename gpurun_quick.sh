python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for bs in 16777216 33554432 67108864; do echo "== batch $bs"; KFRT_BATCH_SLOTS=$bs python tools/counters.py million 0 0 32 2>&1 | tail -2; done
