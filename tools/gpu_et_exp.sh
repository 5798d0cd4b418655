#!/bin/bash
mkdir -p gpurun_out
( for e in 8 24 8 24; do
  echo "== exp $e" 
  SHORT=1 DBG_FLAGS=$((e << 20)) timeout 120 python tools/et_timeline.py 2>&1 | tail -1
done ) > gpurun_out/r02_et_experiments14.txt 2>&1
cat gpurun_out/r02_et_experiments14.txt
