#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_overlap -s 1 -c 1 -o gpurun_out/prof_overlap -f python tools/bench_unitig.py --reads 500000 --err 0.0 --no-ref > gpurun_out/prof_overlap.log 2>&1
ncu -i gpurun_out/prof_overlap.ncu-rep --page raw --csv > gpurun_out/prof_overlap_raw.csv 2>/dev/null
tail -2 gpurun_out/prof_overlap.log
