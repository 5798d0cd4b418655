#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_ncu_list.sh > /dev/null 2>&1
: > gpurun_out/other_workloads.log
for wl in cfg1_monomer64 cfg3_denovo256 cfg4_pmhc800 cfg5_n1024; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> gpurun_out/other_workloads.log
done
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.log
