#!/bin/bash
# 2-GPU sanity of the final build: NCCL determinism test + the two multi-GPU bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/test_gpu_multi_2gpu.log 2>&1; echo "multi test exit $?"; tail -2 gpurun_out/test_gpu_multi_2gpu.log
for c in e2e videos; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 5 --warmup 3 > gpurun_out/bench_${c}_2gpu.json 2> gpurun_out/bench_${c}_2gpu.err; echo "bench $c 2gpu exit $?"
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_${c}_2gpu.json') if l.startswith('{')][-1]); print('$c', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_reference_2gpu.json 2> gpurun_out/bench_reference_2gpu.err; echo "reference arm under torchrun exit $?"; grep -c '^{' gpurun_out/bench_reference_2gpu.json
