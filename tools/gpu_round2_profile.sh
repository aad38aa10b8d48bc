#!/bin/bash
# Round-2 evidence run on one B200: launch list with DRAM bytes of the bench (cheap metrics pass), one full-set capture of the
# three hot kernels, BCR throughput.  Keeps gpurun_out/ small (no source import).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "bcr or encoder" 2>&1 | tail -3
python tools/bench_bcr.py --reads 10000000 --len 150 --ref-reads 0 > gpurun_out/r02_bcr_10M.json 2> gpurun_out/r02_bcr_10M.err; cat gpurun_out/r02_bcr_10M.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_bcr_merge --launch-skip 120 -c 8 --csv \
    --log-file gpurun_out/r02_bcr_merge_ncu.csv python tools/bench_bcr.py --reads 5000000 --len 150 --ref-reads 0 > /dev/null 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-unitig-noisy > gpurun_out/r02_launches.log 2>&1
ncu --set full --clock-control none -k regex:"k_ov_chain|k_ov_nei" -c 2 -o gpurun_out/r02_ov python tools/bench_unitig.py --reads 10000000 --no-ref > gpurun_out/ncu_ov.log 2>&1
ncu --set full --clock-control none --kernel-name-base mangled -k regex:k_smemIjLb1 --launch-skip 6 -c 1 -o gpurun_out/r02_smem_hbm \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-unitig-noisy > gpurun_out/ncu_smem.log 2>&1
ls -la gpurun_out | head -30; du -sh gpurun_out
