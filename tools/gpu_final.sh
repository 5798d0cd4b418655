#!/bin/bash
# end-of-round measurement set: launch lists (cold + warm), default bench line
mkdir -p gpurun_out
bash tools/gpu_ncu_list.sh > /dev/null 2>&1
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.log
