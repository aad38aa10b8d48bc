#!/bin/bash
# the default bench line (SMEM + unitig leg) and the reference arm, timed by the wall clock as the driver does
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/bench_full.json 2> gpurun_out/bench_full.log
tail -4 gpurun_out/bench_full.log; cat gpurun_out/bench_full.json
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log
tail -4 gpurun_out/bench_ref.log; cat gpurun_out/bench_ref.json
