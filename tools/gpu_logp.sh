#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -s -k "option_variants" 2>&1 | tail -15 | tee gpurun_out/pytest_opts.log
