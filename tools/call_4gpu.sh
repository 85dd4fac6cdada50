#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python -c "
import json
s=open('gpurun_out/bench_n4.json').read().strip().splitlines(); print(len(s),'stdout lines'); d=json.loads(s[-1]); print(d['value'], d['e2e']['value'], d['nvlink'])"
tail -3 gpurun_out/bench_n4.err
timeout 120 python - <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe(); g = torch.cuda.device_count(); S = 256 << 20
for name, v in (("push_tma", 2), ("push_sync", 7)):
    r = p.nvlink_a2a(list(range(g)), bytes_per_pair=S, mode=0, warmup=2, reps=8, variant=v)
    print(name, f"ms={r.ms_median:.4f} per_dir={(g-1)*S/r.ms_median/1e6:.1f} own={[round(x) for x in r.egress_gbs[:g]]} verified={r.verified}", flush=True)
PY
