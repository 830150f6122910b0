#!/bin/bash
# scratch GPU round trip: size-specialised tail kernel: parity + A/B timing
mkdir -p gpurun_out
for f in test_gpu_pyramid test_gpu_preproc; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x > gpurun_out/t_$f.log 2>&1; echo "$f exit $?"; tail -3 gpurun_out/t_$f.log
done
for v in 0 1 0 1; do
  MIMAMO_TAIL_GENERIC=$v timeout 300 python bench.py --config e2e --quick --steps 8 --warmup 3 > gpurun_out/exp_e2e_generic$v.json 2> gpurun_out/exp_e2e_generic$v.err; echo "e2e generic=$v exit $?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/exp_e2e_generic$v.json') if l.startswith('{')][-1]); print('generic=$v', d['ms_per_step'], d['value'], d.get('stage_ms'))"
done
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pyramid.py -q -m gpu -x -k "size_specialised" > gpurun_out/san6_memcheck_special.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san6_memcheck_special.log | tail -3
