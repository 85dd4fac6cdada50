#!/bin/bash
mkdir -p gpurun_out
timeout 200 python - <<'PY' 2>&1 | tee gpurun_out/a2a_sync_relaxed_g8.txt
import os, sys
sys.path.insert(0, ".")
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe(); g = torch.cuda.device_count()
for S in (8 << 20, 32 << 20, 64 << 20, 128 << 20, 256 << 20, 1 << 30):
    r = p.nvlink_a2a(list(range(g)), bytes_per_pair=S, mode=0, warmup=2, reps=8, variant=2)
    print(f"S={S>>20}MiB push_tma per_dir={(g-1)*S/r.ms_median/1e6:.1f} ms={r.ms_median:.4f}", flush=True)
    for every in (1, 2, 3, 7):
        os.environ["B200PROBE_A2A_SYNC_EVERY"] = str(every)
        r = p.nvlink_a2a(list(range(g)), bytes_per_pair=S, mode=0, warmup=2, reps=8, variant=7)
        print(f"S={S>>20}MiB push_sync every={every} per_dir={(g-1)*S/r.ms_median/1e6:.1f} ms={r.ms_median:.4f} own={[round(x) for x in r.egress_gbs[:g]]} verified={r.verified}", flush=True)
    os.environ.pop("B200PROBE_A2A_SYNC_EVERY")
PY
timeout 200 python -m pytest tests/test_gpu_a2a.py -m gpu -x -q -k "sync or auto" 2>&1 | tail -3
