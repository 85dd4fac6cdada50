"""Profile target: the NVLink exchange kernels on all visible GPUs.  A2A_PROF_VARIANT picks the schedule
(2 PUSH_TMA concurrent, 6 PUSH_STAGGER = the PUSH_SYNC kernel without its barrier: ncu serialises kernels, so a
cross-GPU barrier could only time out under it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe(); g = torch.cuda.device_count()
os.environ["B200PROBE_A2A_NO_GATE"] = "1"       # the start gate spins on the host: under ncu's serial replay it would be profiled, not the exchange
r = p.nvlink_a2a(list(range(g)), bytes_per_pair=256 << 20, mode=0, warmup=0, reps=1, variant=int(os.environ.get("A2A_PROF_VARIANT", "2")), verify=0)
print(r.ms_median)
