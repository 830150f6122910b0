#!/bin/bash
# One GPU round trip (used through `gpurun -- ./run_gpu_round.sh`): GPU suites, the bench lines, optionally the ncu
# launch list of one quick step (NCU_LIST=1) and a --set full capture (NCU_FULL=<kernel regex>; keep the .ncu-rep small:
# gpurun only copies back 64 MiB -- summarise on the box with tools_ncu_summary.py for anything bigger).
# TESTS="..." selects test files ("none" skips), CONFIGS="e2e pyramid224 resnet512 videos" the bench configurations,
# BENCH_ARGS extra bench.py flags.
mkdir -p gpurun_out
if [ "${TESTS}" != "none" ]; then
for f in ${TESTS:-test_gpu_preproc test_gpu_pyramid test_gpu_nets test_gpu_conv test_gpu_multi}; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?"
  grep -E "passed|failed|Error|error|max\|err" gpurun_out/$f.log | tail -${TAIL:-6}
done
fi
for c in ${CONFIGS:-e2e}; do
  timeout 900 python bench.py --config $c --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
  tail -3 gpurun_out/bench_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$c.json'))
r=d['roofline']
print('$c: ms/step %.2f'%d['ms_per_step'], 'value %.0f %s'%(d['value'], d['unit']), 'e2e %.0f (%.2f ms)'%(d['e2e']['value'], d['e2e']['ms_per_step']),
      'roofline %s %.1f %s frac %.3f'%(r['bound'], r['achieved'], r['unit'], r['frac']), d.get('stage_ms',''), d.get('pyramid_stage_hbm',{}).get('frac',''),
      d.get('sliding_windows',{}).get('ms_per_step',''), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('kind'))
PY
done
if [ -n "$NCU_LIST" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_LIST_SKIP:-0} -c ${NCU_COUNT:-90} --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 1 --warmup 0 ${NCU_BENCH_ARGS:-} > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
fi
if [ -n "$NCU_FULL" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$NCU_FULL" -s ${NCU_SKIP:-0} -c ${NCU_FULL_COUNT:-3} -o gpurun_out/prof_full -f python bench.py --quick --steps 1 --warmup 0 ${NCU_BENCH_ARGS:-} > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
  tail -2 gpurun_out/ncu_full.log
fi
