#!/bin/bash
# One GPU round trip (used through `gpurun -- ./run_gpu_round.sh`): GPU suites, the bench line, optionally the ncu
# launch list of one quick step (NCU_LIST=1) and a --set full capture (NCU_FULL=<kernel regex>; keep the .ncu-rep small:
# gpurun only copies back 64 MiB -- summarise on the box with tools_ncu_summary.py for anything bigger).
# TESTS="..." selects test files, BENCH_ARGS extra bench.py flags.
mkdir -p gpurun_out
for f in ${TESTS:-test_gpu_preproc test_gpu_pyramid test_gpu_nets test_gpu_conv}; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?"
  grep -E "passed|failed|Error|error|max\|err" gpurun_out/$f.log | tail -6
done
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
print('ms/step %.2f'%d['ms_per_step'], 'value %.0f'%d['value'], 'e2e %.0f (%.2f ms)'%(d['e2e']['value'], d['e2e']['ms_per_step']), d['stage_ms'], 'gemm TF %.0f frac %.3f'%(d['roofline']['achieved'], d['roofline']['frac']), d.get('float_inputs'), d.get('cpu_baseline',{}).get('value'))
PY
if [ -n "$NCU_LIST" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c ${NCU_COUNT:-90} --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
fi
if [ -n "$NCU_FULL" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$NCU_FULL" -s ${NCU_SKIP:-0} -c ${NCU_FULL_COUNT:-3} -o gpurun_out/prof_full -f python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
  tail -2 gpurun_out/ncu_full.log
fi
