#!/bin/bash
set -x
mkdir -p gpurun_out
CMD="python bench.py --reads 1000000 --genome 100000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_smem -s 1 -c 1 -o gpurun_out/prof_smem_r1 -f $CMD > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_full.log
ls -la gpurun_out
