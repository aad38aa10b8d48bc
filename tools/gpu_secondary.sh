#!/bin/bash
# secondary workloads of SURVEY.md section 8 on one GPU, each next to the compiled reference; JSON lines -> gpurun_out/secondary.jsonl
mkdir -p gpurun_out; : > gpurun_out/secondary.jsonl
nproc; free -g | head -2
tag() { tail -1 | sed "s/^{/{\"bench\":\"$1\", /" >> gpurun_out/secondary.jsonl; }
python tools/bench_unitig.py --reads 1000000 --err 0.0 2>/dev/null | tag unitig_errfree_1M
python tools/bench_unitig.py --reads 1000000 --err 0.01 2>/dev/null | tag unitig_err1pct_1M
python tools/bench_unitig.py --reads 10000000 --err 0.0 --no-ref 2>/dev/null | tag unitig_errfree_10M
python tools/bench_unitig.py --reads 10000000 --err 0.01 --no-ref 2>/dev/null | tag unitig_err1pct_10M
python tools/bench_encode.py --reads 10000000 2>/dev/null | tail -1 >> gpurun_out/secondary.jsonl
python tools/bench_bcr.py --reads 5000000 --len 150 --ref-reads 300000 --check --fmd 2>/dev/null | tag bcr_5Mx150
python tools/bench_bcr.py --reads 25000000 --len 150 --ref-reads 0 --fmd 2>/dev/null | tag bcr_25Mx150
python tools/bench_ec.py --reads 1000000 2>/dev/null | tag ec_collect_1M
cat gpurun_out/secondary.jsonl
