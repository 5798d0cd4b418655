#!/bin/bash
# node chains: parity (whole GPU suite), then same-box A/B of the step time with and without the chain kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -v "^$" > gpurun_out/pytest_gpu_full.log
tail -30 gpurun_out/pytest_gpu_full.log
for v in 1 0 1 0; do
  FDPT_OPT_6=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chain=$v', 'ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'shares', {k: round(v,3) for k,v in d['time_shares_of_forward'].items()})" | tee -a gpurun_out/ab_chain.log
done
FDPT_OPT_6=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1_monomer64 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/ab_chain.log
FDPT_OPT_6=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1_monomer64 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/ab_chain.log
