#!/bin/bash
mkdir -p gpurun_out
SHORT=1 ncu --set full --import-source on --clock-control none -k regex:et_fused_kernel -s 1 -c 1 -o gpurun_out/et_ncu -f python tools/et_timeline.py > gpurun_out/et_ncu.log 2>&1
ls -la gpurun_out/et_ncu.ncu-rep
tail -3 gpurun_out/et_ncu.log
