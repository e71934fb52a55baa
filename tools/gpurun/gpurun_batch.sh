mkdir -p gpurun_out; : > gpurun_out/batch.log
for b in 33554432 67108864 134217728; do
  echo "== KFRT_BATCH_SLOTS=$b" >> gpurun_out/batch.log
  KFRT_BATCH_SLOTS=$b python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['alu'], d['roofline']['stages_ms_per_step'])" >> gpurun_out/batch.log
done
cat gpurun_out/batch.log
