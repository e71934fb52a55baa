#!/bin/bash
# tools/scale_r2.sh -- multi-GPU evidence of round 2 (run under `gpurun --gpus 8`): the N > 1 product tests,
# the spp-sharded 4K frame and the camera-batch shard (config 5) at 1 / 2 / 4 / 8 GPUs.  One JSON line per run
# in gpurun_out/scale_r2_<mode>_n<N>.json.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
echo "GPUs: $G"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
run() {  # mode N extra...
  local mode=$1 n=$2; shift 2
  local out=gpurun_out/scale_r2_${mode}_n${n}.json
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out 2> gpurun_out/scale_r2_${mode}_n${n}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out 2> gpurun_out/scale_r2_${mode}_n${n}.err
  fi
  python - "$out" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f Mrays/s  %.2f ms/step  e2e %.1f Mrays/s (%.2f ms)  check %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("frame_check")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for n in 1 2 4 8; do
  [ $n -le $G ] || continue
  run 4k $n --width 3840 --height 2160
  run cameras $n --mode cameras
  run 1080p $n
done
