#!/bin/bash
mkdir -p gpurun_out
NO_TIMELINE=1 timeout 300 python tools/et_pair_check.py 2>&1 | tail -60 | tee gpurun_out/et_pair.log
