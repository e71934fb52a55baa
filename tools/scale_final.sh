#!/bin/bash
# tools/scale_final.sh N -- the three bench modes at N GPUs (run under `gpurun --gpus N`); one JSON line per run
# in gpurun_out/r2_scale_<mode>_n<N>.json.  N = 8 also runs the world-8 product tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$1
[ "$n" = 8 ] && timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
run() {
  local mode=$1; shift
  local out=gpurun_out/r2_scale_${mode}_n${n}.json
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out 2> gpurun_out/r2_scale_${mode}_n${n}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out 2> gpurun_out/r2_scale_${mode}_n${n}.err
  fi
  python - "$out" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d["e2e"]
    print(sys.argv[1], "value %.1f  %.2f ms | e2e %.1f (%.2f ms) by-value %s | check %s" % (
        d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e.get("by_value", {}).get("ms_per_step"), d.get("frame_check")))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run 1080p
run 4k --width 3840 --height 2160
run cameras --mode cameras
