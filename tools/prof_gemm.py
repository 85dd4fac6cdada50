"""ncu target: our CTA-pair GEMM and cuBLAS (torch.matmul) on the same 8192^3 U(-1,1) operands, two launches each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from k3s_nvidia_b200.probe import Probe

p = Probe()
M = N = K = 8192
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
A = torch.empty(M * K, dtype=torch.int16, device=dev)
B = torch.empty(N * K, dtype=torch.int16, device=dev)
Cm = torch.empty(M * N, dtype=torch.int16, device=dev)
Cl = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
p._check(p.lib.b200probe_gemm_fill(0, A.data_ptr(), M * K, 0xB200, 2, st), "fill")
p._check(p.lib.b200probe_gemm_fill(0, B.data_ptr(), N * K, 0xB200, 3, st), "fill")
for _ in range(2):
    p._check(p.lib.b200probe_gemm_launch(0, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, st), "gemm_launch")
    torch.matmul(A.view(torch.bfloat16).view(M, K), B.view(torch.bfloat16).view(N, K).t(), out=Cl)
torch.cuda.synchronize()
print("identical", bool((Cm.view(torch.bfloat16).view(M, N) == Cl).all().item()))
