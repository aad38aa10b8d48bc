#!/bin/bash
# multi-GPU unitig only (the driver's scaling run covers the SMEM line): 10 M reads error-free and 1 M reads with errors on N GPUs
N=${1:-8}
mkdir -p gpurun_out
: > gpurun_out/unitig_n$N.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/unitig_multi.py --reads 10000000 --check 2> gpurun_out/unitig_n$N.log | tee -a gpurun_out/unitig_n$N.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/unitig_multi.py --reads 1000000 --err 0.01 --check 2>> gpurun_out/unitig_n$N.log | tee -a gpurun_out/unitig_n$N.jsonl
tail -2 gpurun_out/unitig_n$N.log
