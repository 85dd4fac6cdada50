#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_a2a.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_a2a.txt
timeout 120 python - <<'PY'
import os, sys
sys.path.insert(0, ".")
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe(); g = torch.cuda.device_count(); S = 256 << 20
for name, v in (("push_tma", 2), ("push_stagger", 6), ("push_sync", 7), ("auto", 0)):
    r = p.nvlink_a2a(list(range(g)), bytes_per_pair=S, mode=0, warmup=2, reps=8, variant=v)
    print(name, f"ms={r.ms_median:.4f} per_dir={(g-1)*S/r.ms_median/1e6:.1f} own={[round(x) for x in r.egress_gbs[:g]]} verified={r.verified}", flush=True)
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['nvlink'])"
