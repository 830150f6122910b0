#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -q -m gpu -x > gpurun_out/test_gpu_nets.log 2>&1; echo "nets exit $?"; tail -3 gpurun_out/test_gpu_nets.log
for c in 32 64 128 256; do
  MIMAMO_RESNET_CHUNK=$c timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench.err; echo "chunk $c exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_chunk$c.json'))
print('chunk $c', 'ms/step %.2f'%d['ms_per_step'], 'e2e ms %.2f'%d['e2e']['ms_per_step'], d['stage_ms'], 'gemm TF %.0f'%d['roofline']['achieved'])
PY
done
