#!/bin/bash
# iteration loop on the GPU box: parity first, then the bench line, then a light ncu pass of k_smem
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.log
tail -2 gpurun_out/bench_iter.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print({k:d[k] for k in ('value','ms_per_step')}, 'k_ms', d['roofline']['kernel_ms_per_step'], 'frac', round(d['roofline']['frac'],4), 'e2e', d.get('e2e',{}).get('value'))
PY
if [ -n "$NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_smem -s 1 -c 1 -o gpurun_out/prof_iter -f python bench.py --reads 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_iter.log 2>&1
ncu -i gpurun_out/prof_iter.ncu-rep --page raw --csv > gpurun_out/prof_iter_raw.csv 2>/dev/null
fi
