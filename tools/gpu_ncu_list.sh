#!/bin/bash
# launch lists of one bench command: cold (ncu default: caches flushed between kernels, the recipe's pass) and warm (--cache-control none)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 260 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup > gpurun_out/ncu_bench_warm.log 2>&1
tail -2 gpurun_out/ncu_bench_warm.log | cut -c1-300
