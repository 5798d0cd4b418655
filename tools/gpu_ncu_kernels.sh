#!/bin/bash
# one --set full capture per hot kernel (bench workload cfg2), reports into gpurun_out/; then the launch list
mkdir -p gpurun_out
for k in et_fused_kernel ipa_core_kernel ee_fused_kernel lin_tc_kernel gemm_tc_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 12 -c 1 -o gpurun_out/r01_$k -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-200
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
