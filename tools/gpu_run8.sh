#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1; done > gpurun_out/bench_repeat.log
