#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/other_workloads2.log
for wl in cfg4_tcrpmhc800 cfg5_sweep1024; do
  timeout 900 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/other_workloads2.log
done
