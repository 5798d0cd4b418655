#!/bin/bash
mkdir -p gpurun_out
python tools/ipa_timeline.py 2>&1 | tail -12
tools/gpu_r2_quick.sh
