"""Profile target: the concurrent NVLink exchange (PUSH_TMA, no step barrier: ncu serialises kernels) on all visible GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe(); g = torch.cuda.device_count()
r = p.nvlink_a2a(list(range(g)), bytes_per_pair=256 << 20, mode=0, warmup=0, reps=1, variant=2, verify=0)
print(r.ms_median)
