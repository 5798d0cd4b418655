#!/bin/bash
# bench.py under torchrun on N GPUs of one box (the driver's SCALE launch line); N from $1
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('N=$N value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), 'extra', json.dumps(d.get('extra'))[:1500])
P
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 --impl reference 2>/dev/null | tail -1 | cut -c1-300
