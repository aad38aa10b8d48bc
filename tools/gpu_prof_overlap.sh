#!/bin/bash
# one full ncu capture of each list-chasing overlap phase (k_ov_lists<.,2> and <.,4>) and each chain phase, on the 1 M read index
mkdir -p gpurun_out
READS=${READS:-1000000}
ncu --set full --clock-control none --import-source on -k regex:k_ov_ -s 5 -c 5 -o gpurun_out/prof_overlap -f python tools/bench_unitig.py --reads $READS --err 0.0 --no-ref > gpurun_out/prof_overlap.log 2>&1
tail -2 gpurun_out/prof_overlap.log
ls -la gpurun_out/prof_overlap.ncu-rep
