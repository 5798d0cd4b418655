#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/bench_repeat.log
for i in 1 2 3 4 5; do timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 >> gpurun_out/bench_repeat.log; done
