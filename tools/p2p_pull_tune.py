"""Does pulling over NVLink (TMA bulk loads from the peer) beat pushing (stores)?  2 GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
ords = (C.c_int * 2)(0, 1)
p._check(p.lib.b200probe_enable_peer_access(ords, 2), "peer")
N = 1 << 30
loc = torch.empty(N, dtype=torch.uint8, device="cuda:0")
rem = torch.empty(N, dtype=torch.uint8, device="cuda:1")
torch.cuda.set_device(0)
st = torch.cuda.current_stream(0).cuda_stream
p.hbm_fill(0, loc.data_ptr(), N, 1, st); torch.cuda.synchronize(0)
def timeit(fn, reps=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts)//2]
rows = []
for name, src, dst in (("push(copy loc->rem)", loc, rem), ("pull(copy rem->loc)", rem, loc)):
    for t in (dict(variant=0), dict(variant=0, stage_bytes=16384, stages=4, warps_per_cta=2), dict(variant=0, stage_bytes=8192, stages=6, warps_per_cta=4),
              dict(variant=0, stage_bytes=32768, stages=3, warps_per_cta=2), dict(variant=0, stage_bytes=16384, stages=6, warps_per_cta=2),
              dict(variant=1, ctas_per_sm=4), dict(variant=1, ctas_per_sm=8)):
        ms = timeit(lambda: p.hbm_copy(0, src.data_ptr(), dst.data_ptr(), N, st, **t))
        rows.append(f"{name:22s} {N/ms/1e6:7.1f} GB/s  {t}")
part = torch.zeros(4, dtype=torch.int64, device="cuda:0")
for t in (dict(variant=0), dict(variant=0, stage_bytes=16384, stages=6, warps_per_cta=2), dict(variant=1, ctas_per_sm=4)):
    ms = timeit(lambda: p.hbm_read(0, rem.data_ptr(), N, part.data_ptr(), st, **t))
    rows.append(f"{'read-only from peer':22s} {N/ms/1e6:7.1f} GB/s  {t}")
    ms = timeit(lambda: p.hbm_fill(0, rem.data_ptr(), N, 5, st, **t))
    rows.append(f"{'write-only to peer':22s} {N/ms/1e6:7.1f} GB/s  {t}")
ms = timeit(lambda: rem.copy_(loc)); rows.append(f"{'torch copy_ loc->rem':22s} {N/ms/1e6:7.1f} GB/s")
ms = timeit(lambda: loc.copy_(rem)); rows.append(f"{'torch copy_ rem->loc':22s} {N/ms/1e6:7.1f} GB/s")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/p2p_pull_tune.txt", "w").write("\n".join(rows) + "\n")
print("\n".join(rows))
