python tools/counters.py million 0 0 8 2>&1 | tail -1
python tools/counters.py spheres 0 0 8 2>&1 | tail -1
