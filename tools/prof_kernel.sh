#!/bin/bash
# tools/prof_kernel.sh NAME KERNEL_REGEX SKIP COUNT [ENV=VALUE ...]: one `ncu --set full` capture of the
# launches [SKIP, SKIP+COUNT) matching KERNEL_REGEX of an 8-spp config 3 frame (run under gpurun);
# the report lands in gpurun_out/NAME.ncu-rep and is read here with tools/ncu_summary.py / ncu_lines.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
name=$1; regex=$2; skip=$3; count=$4; shift 4
env "$@" ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c "$count" -f \
  -o gpurun_out/$name python tools/counters.py ${SCENE:-million} ${W:-1920} ${H:-1080} ${SPP:-8} > gpurun_out/$name.log 2>&1
tail -3 gpurun_out/$name.log
