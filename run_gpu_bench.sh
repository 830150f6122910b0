#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py -q -m gpu -x -s > gpurun_out/test_gpu_nets.log 2>&1; echo "nets exit $?"; grep -E "resnet50|passed|failed|Error" gpurun_out/test_gpu_nets.log | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('ms/step %.2f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f (%.2f ms)'%(d['e2e']['value'], d['e2e']['ms_per_step']), d['stage_ms'], 'gemm TF %.0f frac %.3f'%(d['roofline']['achieved'], d['roofline']['frac']))
PY
tail -3 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 70 --csv --log-file gpurun_out/launches_r1.csv python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
