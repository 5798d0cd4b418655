#!/bin/bash
# warm launch list (per-kernel device times, serialised) of the default bench step
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 300 --csv --log-file gpurun_out/r02_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup --no-e2e --no-extra > gpurun_out/ncu_bench_warm.log 2>&1
tail -2 gpurun_out/ncu_bench_warm.log | cut -c1-200
python tools/summarize_launches.py gpurun_out/r02_launches_warm.csv list > gpurun_out/r02_launches_warm_summary.txt
head -45 gpurun_out/r02_launches_warm_summary.txt
