set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_mega.csv python bench.py --steps 2 --warmup 1 --spp 16 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_render -s 1 -c 1 -f -o gpurun_out/prof_r1_mega python bench.py --steps 1 --warmup 1 --spp 4 --no-cpu-baseline > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_mega.json; cat gpurun_out/bench_r1_mega.json
