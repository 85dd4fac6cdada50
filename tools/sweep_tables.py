"""BASELINE configs 2 and 3 as tables: the HBM sweep 1 MiB-1 GiB x {read, write, copy} on GPU 0 and the NVLink
all-to-all sweep 1 MiB-1 GiB per pair over all visible GPUs (AUTO schedule), both verified by the library."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
g = torch.cuda.device_count()
os.makedirs("gpurun_out", exist_ok=True)
lines = ["# HBM sweep, GPU 0: median GB/s over 20 reps after 3 warm-ups, CUDA events; every point verified on the device",
         f"{'bytes':>12} {'read':>9} {'write':>9} {'copy':>9}  cache-resident"]
pts = p.hbm_sweep(0)
by = {}
for q in pts:
    by.setdefault(q.bytes, {})[q.mode] = q
for b in sorted(by):
    r = by[b]
    assert all(x.verified == 1 for x in r.values())
    lines.append(f"{b:>12} {r['read'].gbs_median:>9.1f} {r['write'].gbs_median:>9.1f} {r['copy'].gbs_median:>9.1f}  {int(r['copy'].cache_resident)}")
open("gpurun_out/hbm_sweep_table.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines), flush=True)
if g >= 2:
    lines = [f"# NVLink all-to-all sweep over {g} GPUs, AUTO schedule, per-direction per-GPU GB/s = (G-1)*S / median exchange time (max over devices)",
             f"{'S per pair':>12} {'ms':>9} {'GB/s/dir':>9} {'min own':>9} {'max own':>9} verified"]
    for lg in range(20, 31):
        S = 1 << lg
        if 2 * g * S > (60 << 30):
            break
        r = p.nvlink_a2a(list(range(g)), bytes_per_pair=S, mode=0, warmup=2, reps=8)
        lines.append(f"{S:>12} {r.ms_median:>9.4f} {(g-1)*S/r.ms_median/1e6:>9.1f} {min(r.egress_gbs[:g]):>9.1f} {max(r.egress_gbs[:g]):>9.1f} {r.verified}")
    open(f"gpurun_out/a2a_sweep_table_g{g}.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines), flush=True)
