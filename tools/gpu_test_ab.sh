#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh "$1"
