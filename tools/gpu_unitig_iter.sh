#!/bin/bash
# unitig iteration loop on the GPU box: parity suite (unitig tests), then phase timings at 1 M (vs the reference) and 10 M reads
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} ) 2>&1 | tail -8 > gpurun_out/u_pytest.log
cat gpurun_out/u_pytest.log
python tools/bench_unitig.py --reads 1000000 --err 0.0 > gpurun_out/u_1M.json 2> gpurun_out/u_1M.log
grep "M::fmg_ov\|M::fmg_uni" gpurun_out/u_1M.log | tail -8; cat gpurun_out/u_1M.json
python tools/bench_unitig.py --reads 1000000 --err 0.01 > gpurun_out/u_1Me.json 2> gpurun_out/u_1Me.log
grep "M::fmg_ov\|M::fmg_uni" gpurun_out/u_1Me.log | tail -8; cat gpurun_out/u_1Me.json
python tools/bench_unitig.py --reads 10000000 --err 0.0 --no-ref > gpurun_out/u_10M.json 2> gpurun_out/u_10M.log
grep "M::fmg_ov\|M::fmg_uni" gpurun_out/u_10M.log | tail -8; cat gpurun_out/u_10M.json
