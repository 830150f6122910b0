#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pyramid.py -q -m gpu -x > gpurun_out/test_gpu_pyramid.log 2>&1; echo "pyramid tests exit $?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1100 --csv --log-file gpurun_out/launches_r1.csv python bench.py --quick --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
