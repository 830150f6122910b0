#!/bin/bash
# scratch GPU round trip: new sub-lattice path parity, head timing, halo timing experiments
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -x > gpurun_out/t_conv.log 2>&1; echo "conv exit $?"; tail -3 gpurun_out/t_conv.log
timeout 600 python -m pytest tests/test_gpu_nets.py -q -m gpu -x > gpurun_out/t_nets.log 2>&1; echo "nets exit $?"; tail -3 gpurun_out/t_nets.log
for v in 1 0; do
  MIMAMO_STRIDED_VIEW=$v timeout 300 python bench.py --config e2e --quick --layers --steps 5 --warmup 3 > gpurun_out/exp_e2e_sv$v.json 2> gpurun_out/exp_e2e_sv$v.err; echo "e2e sv=$v exit $?"
  python tools_layers.py gpurun_out/exp_e2e_sv$v.err 2>&1 | tail -2
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/exp_e2e_sv$v.json') if l.startswith('{')][-1]); print('sv=$v', d['ms_per_step'], d['value'], d.get('stage_ms'))"
done
for dbg in 0 1 4 8 12 13 2; do
  MIMAMO_DEBUG=$dbg timeout 300 python bench.py --config resnet512 --quick --layers --steps 3 --warmup 2 > gpurun_out/exp_dbg$dbg.json 2> gpurun_out/exp_dbg$dbg.err; echo "dbg=$dbg exit $?"
  python tools_layers.py gpurun_out/exp_dbg$dbg.err > gpurun_out/exp_dbg$dbg.txt 2>&1
  grep -E "conv1|s2b2 3x3|s3b2 3x3|s4b2 3x3|s5b2 3x3|s2b1 reduce|s4b2 incr|s5b2 incr|per 512" gpurun_out/exp_dbg$dbg.txt | tr '\n' ';'; echo
done
