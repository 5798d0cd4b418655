#!/bin/bash
mkdir -p gpurun_out
python tools/ee_timeline.py 2>&1 | tail -16
tools/gpu_r2_quick.sh
