#!/bin/bash
mkdir -p gpurun_out
MIMAMO_RESNET_CHUNK=128 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 0 -c 1 -o gpurun_out/prof_halo_r1 -f python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
tail -2 gpurun_out/ncu_full.log
