#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python -c "
import json
s=open('gpurun_out/bench_n8.json').read().strip().splitlines(); print(len(s),'stdout line(s)'); d=json.loads(s[-1]); print(d['value'], d['e2e']['value'], d['nvlink'])"
