#!/bin/bash
# round-2 evidence run: full GPU test log, default bench line (with e2e, extras, eager-GPU and CPU baselines), cfg1 latency line,
# reference arm, and the two Linear kernels' ncu captures (demangled names so that template arguments can be matched)
mkdir -p gpurun_out /tmp/ncu
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/r02_pytest_gpu.log
tail -4 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/r02_bench_default.json
tail -3 gpurun_out/bench_err.log
timeout 300 python bench.py --steps 20 --warmup 5 --workload cfg1_monomer64 --no-extra 2>/dev/null | tail -1 > gpurun_out/r02_bench_cfg1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference.json
python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
for k in ('value','ms_per_step','gpu_launches','e2e','e2e_reference_rng_stream','extra','roofline','roofline_ipa','roofline_ipa_core_kernel','cpu_baseline','eager_gpu_baseline','clocks'):
    print(k, json.dumps(d.get(k))[:700])
c=json.load(open('gpurun_out/r02_bench_cfg1.json')); print('cfg1 ms/step', c['ms_per_step'], 'e2e', c.get('e2e',{}).get('value'))
P
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-spinup --no-e2e --no-extra"
cap() {
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$2" -s $3 -c 1 -o /tmp/ncu/$1 -f $B > /tmp/ncu/$1.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$1.ncu-rep > gpurun_out/r02_ncu_full_$1.txt 2>&1
  grep -E "Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |sm__pipe_tensor_cycles_active" gpurun_out/r02_ncu_full_$1.txt | cut -c1-170
}
cap lin_tc_ipa_projection "lin_tc_kernel<.*1>" 4
cap lin_tcw_kernel "lin_tcw_kernel" 40
cap ipa_opt_img_kernel "ipa_opt_img_kernel" 4
du -sh gpurun_out
