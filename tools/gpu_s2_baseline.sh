#!/bin/bash
# session baseline: parity suite, then the unitig path with its phase timings at 1 M and 10 M reads
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8 > gpurun_out/s2_pytest.log
cat gpurun_out/s2_pytest.log
python tools/bench_unitig.py --reads 1000000 --err 0.0 --no-ref > gpurun_out/s2_unitig_1M.json 2> gpurun_out/s2_unitig_1M.log
grep "M::" gpurun_out/s2_unitig_1M.log | tail -12; cat gpurun_out/s2_unitig_1M.json
python tools/bench_unitig.py --reads 10000000 --err 0.0 --no-ref > gpurun_out/s2_unitig_10M.json 2> gpurun_out/s2_unitig_10M.log
grep "M::" gpurun_out/s2_unitig_10M.log | tail -30; cat gpurun_out/s2_unitig_10M.json
