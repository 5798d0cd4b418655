#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -v "^$" > gpurun_out/pytest_gpu_full.log
tail -6 gpurun_out/pytest_gpu_full.log
rm -f gpurun_out/ab_tfimg.log
for v in 1 0 1 0; do
  FDPT_OPT_8=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tf_img=$v', 'ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'seq_tfmr share', round(d['time_shares_of_forward']['seq_tfmr'],4))" | tee -a gpurun_out/ab_tfimg.log
done
