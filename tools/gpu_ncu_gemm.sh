#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/gemm_proj -f python tools/one_gemm.py 16384 6816 256 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/gemm_small -f python tools/one_gemm.py 2800 256 256 >> gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
