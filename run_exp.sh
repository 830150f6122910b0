#!/bin/bash
# profiling round trip of the final build: launch list of one step, --set full of the specialised tail kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final2.csv python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"; grep -c "gpu__time_duration" gpurun_out/launches_r2_final2.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phase_tail_map -s 2 -c 2 -o gpurun_out/prof_tail2 -f python bench.py --quick --steps 1 --warmup 0 > gpurun_out/ncu_tail2.log 2>&1; echo "ncu tail exit $?"
python tools_ncu_summary.py gpurun_out/prof_tail2.ncu-rep > gpurun_out/prof_tail2_summary.txt 2>&1; tail -40 gpurun_out/prof_tail2_summary.txt
ls -la gpurun_out/prof_tail2.ncu-rep
