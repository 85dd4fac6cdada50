#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/probe_round.py 2>&1 | tail -3
timeout 300 python tools/probe_round.py --keep 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 200 --warmup 5 --no-gemm > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('probe_round')))
PY
