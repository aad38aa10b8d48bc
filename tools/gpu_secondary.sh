#!/bin/bash
# secondary workloads of SURVEY.md section 8 on one GPU, each next to the compiled reference; JSON lines -> gpurun_out/secondary.jsonl
mkdir -p gpurun_out; : > gpurun_out/secondary.jsonl
python tools/bench_unitig.py --reads 1000000 --err 0.0 2>/dev/null | tail -1 | sed 's/^/{"bench":"unitig_errfree",/; s/{"reads/"reads/' >> gpurun_out/secondary.jsonl
python tools/bench_unitig.py --reads 1000000 --err 0.01 2>/dev/null | tail -1 | sed 's/^/{"bench":"unitig_err1pct",/; s/{"reads/"reads/' >> gpurun_out/secondary.jsonl
python tools/bench_unitig.py --reads 5000000 --err 0.0 --no-ref 2>/dev/null | tail -1 | sed 's/^/{"bench":"unitig_5M",/; s/{"reads/"reads/' >> gpurun_out/secondary.jsonl
python tools/bench_bcr.py --reads 5000000 --len 150 --ref-reads 300000 --check 2>/dev/null | tail -1 | sed 's/^/{"bench":"bcr_5Mx150",/; s/{"reads/"reads/' >> gpurun_out/secondary.jsonl
python tools/bench_bcr.py --reads 25000000 --len 150 --ref-reads 0 2>/dev/null | tail -1 | sed 's/^/{"bench":"bcr_25Mx150",/; s/{"reads/"reads/' >> gpurun_out/secondary.jsonl
python tools/bench_ec.py --reads 1000000 2>/dev/null | tail -1 | sed 's/^/{"bench":"ec_collect_1M",/; s/{"reads/"reads/' >> gpurun_out/secondary.jsonl
cat gpurun_out/secondary.jsonl
