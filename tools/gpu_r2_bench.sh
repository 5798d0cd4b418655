#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" > gpurun_out/pytest_gpu_full.log
tail -8 gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_r2.json
tail -5 gpurun_out/bench_err.log
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r2.json'))
for k in ('value','ms_per_step','gpu_launches','e2e','e2e_reference_rng_stream','extra','roofline_ipa','roofline_ipa_core_kernel','cpu_baseline','clocks','dtype'):
    print(k, json.dumps(d.get(k))[:600])
P
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-900
