#!/bin/bash
mkdir -p gpurun_out
for w in cfg1_monomer64 cfg3_denovo256 cfg4_tcrpmhc800 cfg5_sweep1024; do
  echo "== $w"; timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-1500
done | tee gpurun_out/other_workloads.log
