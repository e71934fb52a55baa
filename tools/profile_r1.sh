#!/bin/bash
# Round-1 ncu evidence for the wavefront path (run under gpurun; results land in gpurun_out/).
#   launches_r1_wf.csv    every stage launch of one 32-spp batch with its device time
#   traffic_r1_wf.csv     dram bytes of every k_wf_trace launch of that batch
#   prof_r1_trace.ncu-rep one --set full capture of k_wf_trace<closest hit> (depth 1, incoherent rays, 8-spp batch)
set -x
mkdir -p gpurun_out
# 32 spp = one full 64 Mi-slot batch: the launches ncu sees have the size of the bench's launches
CMD="python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu-baseline"
CMD8="python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_wf_|k_resolve' -c 200 --csv \
    --log-file gpurun_out/launches_r1_wf.csv $CMD > gpurun_out/bench_under_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_wf_trace -c 18 --csv --log-file gpurun_out/traffic_r1_wf.csv $CMD >> gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 2 -c 2 -f \
    -o gpurun_out/prof_r1_trace $CMD8 >> gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log
