"""Profile target: one launch each of the HBM read / write / copy ring kernels and the verify kernel at 1 GiB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k3s_nvidia_b200.probe import Probe
from k3s_nvidia_b200 import _lib as L
p = Probe()
pts = p.hbm_sweep(0, min_bytes=1 << 30, max_bytes=1 << 30, modes=L.HBM_READ | L.HBM_WRITE | L.HBM_COPY, warmup=0, reps=1, verify=1)
print([(q.mode, round(q.gbs_median)) for q in pts])
