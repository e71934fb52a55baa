#!/bin/bash
# tools/ab_env.sh OUT "ENV=V ..." ...: stage times of an 8-spp config 3 frame under each environment setting
# (run under gpurun); the log lands in gpurun_out/OUT.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/$1; shift
: > $out
for e in "$@"; do
  echo "=== $e" >> $out
  env $e timeout 180 python tools/counters.py ${SCENE:-million} ${W:-1920} ${H:-1080} ${SPP:-8} 2>&1 | grep -E "per ray|frame|stages|rror" >> $out
done
cat $out
