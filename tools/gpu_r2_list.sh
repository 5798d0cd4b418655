#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup --no-e2e --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 300 --csv --log-file /tmp/ncu/warm.csv $B > /tmp/ncu/warm.log 2>&1
python tools/summarize_launches.py /tmp/ncu/warm.csv list > gpurun_out/r02_launches_warm_summary.txt
head -40 gpurun_out/r02_launches_warm_summary.txt
