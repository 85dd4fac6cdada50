#!/bin/bash
mkdir -p gpurun_out
timeout 60 python - <<'PY' 2>&1 | tee gpurun_out/a2a_pair_matrix_g4.txt
import sys, subprocess
sys.path.insert(0, ".")
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe(); g = torch.cuda.device_count()
for S in (256 << 20, 8 << 20):
    r = p.nvlink_a2a(list(range(g)), bytes_per_pair=S, mode=0, warmup=2, reps=6, variant=7)
    print(f"S={S>>20}MiB PUSH_SYNC per_dir={(g-1)*S/r.ms_median/1e6:.1f} ms={r.ms_median:.4f} verified={r.verified} min_pair={r.min_pair_gbs:.1f} max_pair={r.max_pair_gbs:.1f}")
    for row in r.pair_gbs:
        print("   ", [round(x, 1) for x in row])
    tot = [sum(S / (x * 1e6) for x in row if x > 0) for row in r.pair_gbs]
    print("    sum of step times per rank (ms):", [round(x, 4) for x in tot], flush=True)
import pytest
sys.exit(pytest.main(["tests/test_gpu_a2a.py", "-m", "gpu", "-x", "-q", "-k", "sync or auto or selectors"]))
PY
