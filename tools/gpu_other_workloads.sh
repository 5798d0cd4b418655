#!/bin/bash
mkdir -p gpurun_out
for w in cfg1_monomer64 cfg3_denovo256 cfg4_tcrpmhc800 cfg5_sweep1024; do
  echo "== $w"; timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
done > gpurun_out/other_workloads.log 2>&1
tail -c 300 gpurun_out/other_workloads.log
