set -x
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
