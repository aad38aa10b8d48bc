#!/bin/bash
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
python bench.py --reads 2000000 --genome 20000000 --steps 3 --warmup 2 --cpu-seconds 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.log
tail -5 gpurun_out/bench_small.log; cat gpurun_out/bench_small.json
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.log
tail -8 gpurun_out/bench_full.log; cat gpurun_out/bench_full.json
