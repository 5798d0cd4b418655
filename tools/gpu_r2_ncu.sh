#!/bin/bash
# one `ncu --set full` capture per hot kernel at cfg2 (bench default) and, for the pair kernels, at cfg3; reports stay on the box
# (/tmp), only their text summaries (tools/ncu_summary.py) and the launch lists come back
mkdir -p gpurun_out /tmp/ncu
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup --no-e2e --no-extra"
cap() {  # name, kernel regex, skip, extra bench args
  ncu --set full --import-source on --clock-control none -k regex:"$2" -s $3 -c 1 -o /tmp/ncu/$1 -f $B $4 > /tmp/ncu/$1.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$1.ncu-rep > gpurun_out/r02_ncu_full_$1.txt 2>&1
  grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |sm__pipe_tensor_cycles_active|gpu__dram_throughput" gpurun_out/r02_ncu_full_$1.txt | awk '{print "'$1'", $1, $2, $3}'
}
cap et_fused_kernel et_fused_kernel 6 ""
cap ipa_core_kernel ipa_core_kernel 6 ""
cap ee_fused_kernel ee_fused_kernel 2 ""
cap gemm_img_kernel gemm_img_kernel 6 ""
cap gemm_tc_kernel gemm_tc_kernel 6 ""
cap softmax_rows_kernel softmax_rows_kernel 6 ""
cap rot_score_kernel rot_score_kernel 2 ""
cap reverse_kernel reverse_kernel 2 ""
cap cfg3_et_fused_kernel et_fused_kernel 4 "--workload cfg3_denovo256"
cap cfg3_ipa_core_kernel ipa_core_kernel 4 "--workload cfg3_denovo256"
cap cfg3_gemm_img_kernel gemm_img_kernel 4 "--workload cfg3_denovo256"
cap cfg4_et_fused_kernel et_fused_kernel 4 "--workload cfg4_tcrpmhc800 --global-batch 8"
cap cfg4_ipa_core_kernel ipa_core_kernel 4 "--workload cfg4_tcrpmhc800 --global-batch 8"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 300 --csv --log-file /tmp/ncu/warm.csv $B > /tmp/ncu/warm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file /tmp/ncu/cold.csv $B > /tmp/ncu/cold.log 2>&1
python tools/summarize_launches.py /tmp/ncu/warm.csv list > gpurun_out/r02_launches_warm_summary.txt
python tools/summarize_launches.py /tmp/ncu/cold.csv > gpurun_out/r02_launches_cold_summary.txt
head -12 gpurun_out/r02_launches_warm_summary.txt
du -sh gpurun_out
