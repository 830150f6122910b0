#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel kills only that file's context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_pyramid test_gpu_conv test_gpu_nets; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
