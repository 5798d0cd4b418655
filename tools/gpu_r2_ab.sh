#!/bin/bash
# generic same-box A/B: bash tools/gpu_r2_ab.sh <option number> <values...>; runs the GPU tests first
mkdir -p gpurun_out
OPT=$1; shift
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -v "^$" > gpurun_out/pytest_gpu_full.log
tail -5 gpurun_out/pytest_gpu_full.log
rm -f gpurun_out/ab_opt$OPT.log
for rep in 1 2; do for v in "$@"; do
  env FDPT_OPT_$OPT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('opt$OPT=$v', 'ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'shares', {k: round(x,3) for k,x in d['time_shares_of_forward'].items()})" | tee -a gpurun_out/ab_opt$OPT.log
done; done
