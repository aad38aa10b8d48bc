#!/bin/bash
# First-contact script for a GPU box: smoke, GPU parity tests, then a short perf probe.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
