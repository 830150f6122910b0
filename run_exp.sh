#!/bin/bash
# scratch GPU round trip: shared blur denominator + 3 CTAs / SM: parity + A/B timing
mkdir -p gpurun_out
for f in test_gpu_pyramid test_gpu_preproc; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x > gpurun_out/t_$f.log 2>&1; echo "$f exit $?"; tail -3 gpurun_out/t_$f.log
done
for v in "3 1" "2 1" "2 0" "3 1" "2 1" "2 0"; do
  set -- $v
  MIMAMO_TAIL_OCC=$1 MIMAMO_TAIL_DEN=$2 timeout 300 python bench.py --config e2e --quick --steps 8 --warmup 3 > gpurun_out/exp_e2e_occ$1_den$2.json 2> gpurun_out/exp_e2e_occ$1_den$2.err; echo "e2e occ=$1 den=$2 exit $?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/exp_e2e_occ$1_den$2.json') if l.startswith('{')][-1]); print('occ=$1 den=$2', d['ms_per_step'], d['value'], d.get('stage_ms'))"
done
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pyramid.py -q -m gpu -x -k "shared_blur" > gpurun_out/san5_memcheck_den.log 2>&1; echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san5_memcheck_den.log | tail -3
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_pyramid.py -q -m gpu -x -k "shared_blur and 48" > gpurun_out/san5_racecheck_den.log 2>&1; echo "racecheck exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san5_racecheck_den.log | tail -3
