#!/bin/bash
# GPU tests (stop at first failure) + two default-bench step times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -v "^$" > gpurun_out/pytest_gpu_full.log
tail -5 gpurun_out/pytest_gpu_full.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'ET ms', round(d['roofline_edge_transition']['avg_ms'],4), 'frac', round(d['roofline_edge_transition']['frac'],3), 'shares', {k: round(x,3) for k,x in d['time_shares_of_forward'].items()})"
done
