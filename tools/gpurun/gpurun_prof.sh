mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 2 -c 1 -f -o gpurun_out/prof_shade python tools/counters.py million 0 0 8 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
