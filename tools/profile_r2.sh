#!/bin/bash
# Round-2 ncu evidence (run under gpurun; results land in gpurun_out/, digests are copied to profiles/).
#   r2_launches.csv        every stage launch of one 32-spp batch of the bench frame with its device time
#   r2_traffic.csv         dram bytes + time of every k_wf_trace / k_wf_shade / k_wf_shadow_resolve launch of that batch
#   r2_trace.ncu-rep       --set full of k_wf_trace (depth-1 closest hit, then its occlusion launch), 8-spp batch
#   r2_shade.ncu-rep       --set full of k_wf_shade (depth 1) and k_wf_shadow_resolve
#   r2_unique_traffic.csv / r2_unique10m_traffic.csv   dram bytes of the traversal launches on the un-instanced stress scenes
set -x
# the structure probe (kfrtSetInstanceSubtrees mode 1) picks the two-level structure for config 3; it is switched
# off here so that the captures below see the kernels of the frame, not those of the probe
export KFRT_INSTANCE_SUBTREES=0
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --spp 32 --no-cpu-baseline"
CMD8="python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_wf_|k_resolve' -c 200 --csv \
    --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_wf_trace|k_wf_shade|k_wf_shadow_resolve' -c 36 --csv --log-file gpurun_out/r2_traffic.csv $CMD >> gpurun_out/r2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 2 -c 2 -f \
    -o gpurun_out/r2_trace $CMD8 >> gpurun_out/r2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_wf_shade|k_wf_shadow_resolve' -s 2 -c 2 -f \
    -o gpurun_out/r2_shade $CMD8 >> gpurun_out/r2_ncu.log 2>&1
for s in unique unique10m; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none \
      -k regex:k_wf_trace -c 18 --csv --log-file gpurun_out/r2_${s}_traffic.csv \
      python bench.py --scene $s --steps 1 --warmup 1 --spp 8 --no-cpu-baseline >> gpurun_out/r2_ncu.log 2>&1
  python tools/counters.py $s 1920 1080 8 > gpurun_out/r2_${s}_counters.txt 2>&1
done
tail -2 gpurun_out/r2_ncu.log
