#!/bin/bash
# Final round-2 validation on one B200: every GPU suite, smoke, the bench lines, the launch list of one step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest gpu exit $?"; tail -2 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench e2e exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_e2e.json 2> gpurun_out/bench_reference_e2e.err; echo "bench reference exit $?"
timeout 600 python bench.py --config resnet512 --steps 10 --warmup 3 > gpurun_out/bench_resnet512.json 2> gpurun_out/bench_resnet512.err; echo "bench resnet512 exit $?"
timeout 600 python bench.py --config pyramid224 --steps 5 --warmup 3 > gpurun_out/bench_pyramid224.json 2> gpurun_out/bench_pyramid224.err; echo "bench pyramid224 exit $?"
timeout 600 python bench.py --config pyramid224 --pyr-height 6 --pyr-levels 1,2,3,4 --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_pyramid224_h6.json 2> gpurun_out/bench_pyramid224_h6.err; echo "bench pyramid224 h6 exit $?"; tail -2 gpurun_out/bench_pyramid224_h6.err
python - <<'PY'
import json
for c in ("e2e", "reference_e2e", "resnet512", "pyramid224", "pyramid224_h6"):
    try:
        d = json.loads([l for l in open('gpurun_out/bench_%s.json' % c) if l.startswith('{')][-1])
        r = d.get('roofline', {})
        print(c, 'ms/step %.2f' % d['ms_per_step'], 'value %.1f' % d['value'], 'e2e %.1f' % d['e2e']['value'], 'frac', r.get('frac'), d.get('stage_ms', ''),
              d.get('sliding_windows', {}).get('ms_per_step', ''), 'cpu', d.get('cpu_baseline', {}).get('value'), d.get('clocks'))
    except Exception as e:
        print(c, 'unreadable', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r2_final2.csv python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
