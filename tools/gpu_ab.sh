#!/bin/bash
# same-box A/B of FDPT_OPT_DEBUG_FLAGS settings: bash tools/gpu_ab.sh "0 256 16384"  -> gpurun_out/ab.log (2 rounds, interleaved)
mkdir -p gpurun_out
: > gpurun_out/ab.log
for round in 1 2; do
  for f in $1; do
    echo -n "flags=$f " >> gpurun_out/ab.log
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --debug-flags $f 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))" >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
