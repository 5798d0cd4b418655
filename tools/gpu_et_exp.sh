#!/bin/bash
mkdir -p gpurun_out
( for e in 8 9 10 11 12; do
  echo "== exp $e" 
  SHORT=1 DBG_FLAGS=$((e << 20)) timeout 120 python tools/et_timeline.py 2>&1 | tail -1
done
echo "== exp 11 timeline"
DBG_FLAGS=$((11 << 20)) timeout 120 python tools/et_timeline.py 2>&1 | tail -37 ) > gpurun_out/r02_et_experiments15.txt 2>&1
cat gpurun_out/r02_et_experiments15.txt
