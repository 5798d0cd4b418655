#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/bench_repeat.log
for i in 1 2 3; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/bench_repeat.log; done
timeout 200 python tools/bench_lin.py 2>&1 | tail -8 > gpurun_out/bench_lin.log
