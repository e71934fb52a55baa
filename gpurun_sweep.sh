for tb in 1 6 12 16 24; do for ib in 1 8 16; do echo "== tri $tb inst $ib"; KFRT_TRI_BATCH=$tb KFRT_INST_BATCH=$ib python tools/counters.py million 0 0 8 2>&1 | tail -4; done; done
