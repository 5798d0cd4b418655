#!/bin/bash
# what the driver runs at round end: GPU tests, smoke(), default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.log
