#!/bin/bash
# N-GPU checks: the bench line under torchrun and the multi-GPU unitig (NCCL all-reduce + all-gather of the record shards, GPU assembly)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log
tail -3 gpurun_out/bench_n$N.log; cat gpurun_out/bench_n$N.json
: > gpurun_out/unitig_n$N.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/unitig_multi.py --reads 1000000 --check 2> gpurun_out/unitig_n$N.log | tee -a gpurun_out/unitig_n$N.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/unitig_multi.py --reads 1000000 --err 0.01 --check 2>> gpurun_out/unitig_n$N.log | tee -a gpurun_out/unitig_n$N.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/unitig_multi.py --reads 10000000 --check 2>> gpurun_out/unitig_n$N.log | tee -a gpurun_out/unitig_n$N.jsonl
tail -3 gpurun_out/unitig_n$N.log
