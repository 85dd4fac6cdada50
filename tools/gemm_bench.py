import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from k3s_nvidia_b200.probe import Probe
p = Probe()
v = os.environ.get("B200PROBE_GEMM_VARIANT", "default")
for s in (2048, 4096, 8192):
    r = p.gemm(0, m=s, n=s, k=s, warmup=3, reps=10)
    print("variant", v, s, round(r.tflops_median, 1), round(r.tflops_best, 1), r.verified)
r = p.gemm(0, warmup=3, reps=5, sustain_seconds=4.0)
print("variant", v, "sustained", round(r.tflops_sustained, 1), "burst", round(r.tflops_best, 1))
import torch
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3): a @ b.T
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b.T; e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("torch.matmul (cuBLAS) same call: best", round(2 * 8192**3 / min(ts) / 1e9, 1), "median", round(2 * 8192**3 / sorted(ts)[5] / 1e9, 1))
