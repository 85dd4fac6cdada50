#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_1.txt
python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe(); n = 1 << 30
buf = torch.empty(n, dtype=torch.uint8, device="cuda:0")
p.hbm_fill(0, buf.data_ptr(), n, 0xB200, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
for _ in range(3): p.hbm_verify(0, buf.data_ptr(), n, 0xB200)
t = time.perf_counter()
for _ in range(50): r = p.hbm_verify(0, buf.data_ptr(), n, 0xB200)
dt = (time.perf_counter() - t) / 50
print(f"hbm_verify 1 GiB (sync call): {dt*1e6:.1f} us -> {n/dt/1e9:.0f} GB/s, bad={r[2]}")
PY
python bench.py --steps 500 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['e2e_hostbuf']['value'])"; tail -3 gpurun_out/bench_n1.err
