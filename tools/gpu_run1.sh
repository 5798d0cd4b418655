#!/bin/bash
# bring-up of gemm_tc + regression + bench
mkdir -p gpurun_out
for c in kmajor mn0 mn1 simt; do
  echo "=== $c"; timeout 120 python tools/bringup_gemm_tc.py $c 2>&1 | tail -12
done > gpurun_out/bringup_gemm_tc.log 2>&1
cat gpurun_out/bringup_gemm_tc.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.log
