#!/bin/bash
mkdir -p gpurun_out
( for e in 0; do
  echo "== exp $e" 
  DBG_FLAGS=$((e << 20)) timeout 120 python tools/et_timeline.py 2>&1 | tail -50
done ) > gpurun_out/r02_et_experiments5.txt 2>&1
tail -3 gpurun_out/r02_et_experiments5.txt
tools/gpu_r2_quick.sh
