#!/bin/bash
# per-step time as a function of the number of timed steps (same box)
mkdir -p gpurun_out
: > gpurun_out/steps_sweep.log
for k in 5 20; do
  echo -n "steps=$k " >> gpurun_out/steps_sweep.log
  timeout 600 python bench.py --steps $k --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['diag']['enqueue_wall_ms'])" >> gpurun_out/steps_sweep.log
done
cat gpurun_out/steps_sweep.log
