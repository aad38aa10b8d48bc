#!/bin/bash
# Evidence for profiles/: the bench line (SMEM + unitig leg) and the reference arm, the ncu launch lists of the same workloads,
# one full capture of k_smem and of the four overlap phase kernels (on the 2x10^9-symbol index: the HBM regime).
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.log
cat gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.log
cat gpurun_out/bench_r1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --unitig-reads 0 > gpurun_out/launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_unitig.csv python tools/bench_unitig.py --reads 10000000 --err 0.0 --no-ref > gpurun_out/launches_unitig.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_smem -s 2 -c 1 -o gpurun_out/prof_k_smem -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --unitig-reads 0 > gpurun_out/prof_k_smem.log 2>&1
ncu -i gpurun_out/prof_k_smem.ncu-rep --page raw --csv > gpurun_out/prof_k_smem_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_ov_ -s 50 -c 5 -o gpurun_out/prof_overlap_10M -f python tools/bench_unitig.py --reads 10000000 --err 0.0 --no-ref > gpurun_out/prof_overlap_10M.log 2>&1
ncu -i gpurun_out/prof_overlap_10M.ncu-rep --page raw --csv > gpurun_out/prof_overlap_10M_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
