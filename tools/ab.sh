#!/bin/bash
# tools/ab.sh OUT VARIANT...: stage times of an 8-spp config 3 frame for every kuafu_b200/lib_VARIANT (run under
# gpurun after tools/variant.sh built them); the log lands in gpurun_out/OUT.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/$1; shift
: > $out
for v in "$@"; do
  echo "=== $v" >> $out
  KFRT_LIB_DIR=kuafu_b200/lib_$v timeout 120 python tools/counters.py ${SCENE:-million} 1920 1080 ${SPP:-8} 2>&1 | grep -E "frame|stages|rror" >> $out
done
cat $out
