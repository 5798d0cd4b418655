#!/bin/bash
mkdir -p gpurun_out
LT_FLAG=524288 python tools/lin_timeline.py 2>&1 | grep -A13 "320x320" 
tools/gpu_r2_quick.sh
