#!/bin/bash
# Evidence for profiles/: the bench line, the ncu launch list of the same command, one full capture of k_smem.
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.log
cat gpurun_out/bench_r1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.log
cat gpurun_out/bench_r1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_smem -s 2 -c 1 -o gpurun_out/prof_k_smem -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_k_smem.log 2>&1
ncu -i gpurun_out/prof_k_smem.ncu-rep --page raw --csv > gpurun_out/prof_k_smem_raw.csv 2>/dev/null
