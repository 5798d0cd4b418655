#!/bin/bash
# round 2, first GPU call: the whole GPU test suite (new parity cases on the bench configurations included) + the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench.log
