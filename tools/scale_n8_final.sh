#!/bin/bash
# tools/scale_n8_final.sh -- the spp-sharded 1080p frame and the camera-sharded config 5 on 8 GPUs with the final
# kernels (run under `gpurun --gpus 8`); one JSON line each in gpurun_out/r2_scale_<mode>_n8_final.json.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  local mode=$1; shift
  local out=gpurun_out/r2_scale_${mode}_n8_final.json
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline "$@" > $out 2> gpurun_out/r2_scale_${mode}_n8_final.err
  python - "$out" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d["e2e"]
    print(sys.argv[1], "value %.1f  %.2f ms | e2e %.1f (%.2f ms) | check %s" % (d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], d.get("frame_check")))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
run 1080p
run cameras --mode cameras
