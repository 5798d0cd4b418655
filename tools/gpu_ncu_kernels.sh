#!/bin/bash
# one --set full capture per hot kernel (bench workload cfg2), reports into gpurun_out/
mkdir -p gpurun_out
for k in et_fused_kernel ipa_core_kernel ee_fused_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 6 -c 1 -o gpurun_out/r01_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
