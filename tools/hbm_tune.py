"""Tuning sweep for the HBM kernels on a real B200: every (variant, stage_bytes, stages, warps,
ctas_per_sm) at one buffer size, timed with CUDA events on the launch stream.  Output: a table
sorted by GB/s (gpurun_out/hbm_tune.txt).  Usage: python tools/hbm_tune.py [bytes_log2=30]"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from k3s_nvidia_b200.probe import Probe

LOG2 = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N = 1 << LOG2
p = Probe()
src = torch.empty(N, dtype=torch.uint8, device="cuda:0")
dst = torch.empty(N, dtype=torch.uint8, device="cuda:0")
part = torch.zeros(4, dtype=torch.int64, device="cuda:0")
st = torch.cuda.current_stream().cuda_stream
p.hbm_fill(0, src.data_ptr(), N, 0xB200, st)
torch.cuda.synchronize()


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


rows = []
combos = []
for sb, stg, w, c in itertools.product([4096, 8192, 16384, 32768], [2, 3, 4, 6, 8], [1, 2, 4, 8], [1, 2]):
    if sb * stg * w * c + 2048 * c > 227 * 1024:
        continue
    if sb * stg * w * c < 48 * 1024:
        continue
    combos.append(dict(variant=0, stage_bytes=sb, stages=stg, warps_per_cta=w, ctas_per_sm=c))
for c in [1, 2, 3, 4, 6, 8]:
    combos.append(dict(variant=1, ctas_per_sm=c))

modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["copy", "read", "write"]
for t in combos:
    for mode in modes:
        try:
            if mode == "copy":
                fn = lambda: p.hbm_copy(0, src.data_ptr(), dst.data_ptr(), N, st, **t)
                alg = 2 * N
            elif mode == "read":
                fn = lambda: p.hbm_read(0, src.data_ptr(), N, part.data_ptr(), st, **t)
                alg = N
            else:
                fn = lambda: p.hbm_fill(0, dst.data_ptr(), N, 0xB200, st, **t)
                alg = N
            med, best = timeit(fn)
            rows.append((alg / med / 1e6, alg / best / 1e6, mode, t))
        except Exception as e:  # noqa: BLE001
            rows.append((0.0, 0.0, mode, dict(t, err=str(e)[:80])))

# library baseline for context: torch copy_ (what MEASURED_PEAKS.json's hbm_gbs is)
a = src.view(torch.bfloat16)
b = dst.view(torch.bfloat16)
med, best = timeit(lambda: b.copy_(a))
rows.append((2 * N / med / 1e6, 2 * N / best / 1e6, "copy", dict(variant="torch.copy_")))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/hbm_tune.txt", "w") as f:
    for mode in modes:
        f.write(f"== {mode} @ 2^{LOG2} B  (GB/s median, best)\n")
        for r in sorted([r for r in rows if r[2] == mode], key=lambda r: -r[0]):
            f.write(f"{r[0]:8.1f} {r[1]:8.1f}  {r[3]}\n")
print(open("gpurun_out/hbm_tune.txt").read()[:6000])
