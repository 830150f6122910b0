#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 2 -c 3 -o gpurun_out/prof_gemm_r1 -f python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
