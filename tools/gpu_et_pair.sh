#!/bin/bash
mkdir -p gpurun_out
ET_PAIR=1 DBG_FLAGS=64 timeout 120 python tools/et_timeline.py 2>&1 | tail -36 > gpurun_out/et_tl_pair_nomma.log
ET_PAIR=1 DBG_FLAGS=128 timeout 120 python tools/et_timeline.py 2>&1 | tail -36 > gpurun_out/et_tl_pair_2xmma.log
