#!/bin/bash
mkdir -p gpurun_out
for c in 512 1024 2048; do
  MIMAMO_RESNET_CHUNK=$c timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench.err; echo "chunk $c exit $?"; tail -2 gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_chunk$c.json'))
print('chunk $c', 'ms/step %.2f'%d['ms_per_step'], 'e2e ms %.2f'%d['e2e']['ms_per_step'], d['stage_ms'], 'gemm TF %.0f'%d['roofline']['achieved'])
PY
done
