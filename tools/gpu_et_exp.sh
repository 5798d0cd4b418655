#!/bin/bash
mkdir -p gpurun_out
( for e in 8; do
  echo "== exp $e" 
  DBG_FLAGS=$((e << 20)) timeout 120 python tools/et_timeline.py 2>&1 | tail -37
done ) > gpurun_out/r02_et_experiments13.txt 2>&1
grep "exp\|period" gpurun_out/r02_et_experiments13.txt
tools/gpu_r2_quick.sh
