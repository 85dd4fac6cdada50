"""Run-to-run spread of the three probe figures (north_star: floating-point probe results stable within 1 % run to run):
10 repeats each of the plugin-facing calls, (max - min) / median.  Writes gpurun_out/stability_g<G>.txt."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from k3s_nvidia_b200 import _lib as L
from k3s_nvidia_b200.probe import Probe

p = Probe()
g = min(torch.cuda.device_count(), 8)
N = int(os.environ.get("STAB_REPEATS", "10"))
lines = [f"# {N} repeats of each plugin-facing probe call; spread = (max - min) / median"]


def report(name, vals, unit):
    sp = (max(vals) - min(vals)) / statistics.median(vals)
    line = f"{name:<44} median {statistics.median(vals):9.1f} {unit}  min {min(vals):9.1f}  max {max(vals):9.1f}  spread {100 * sp:5.2f} %"
    lines.append(line)
    print(line, flush=True)


kw = dict(min_bytes=1 << 30, max_bytes=1 << 30, warmup=3, reps=20, verify=1)
ONLY_NVLINK = bool(os.environ.get("STAB_ONLY_NVLINK"))
for mode, name in (() if ONLY_NVLINK else ((L.HBM_COPY, "copy"), (L.HBM_READ, "read"), (L.HBM_WRITE, "write"))):
    p.hbm_sweep(0, modes=mode, **kw)
    report(f"HBM {name} 1 GiB (median of 20 reps per call)", [p.hbm_sweep(0, modes=mode, **kw)[0].gbs_median for _ in range(N)], "GB/s")
for cls, name in (() if ONLY_NVLINK else ((L.GEMM_EXACT, "k/128 operands"), (L.GEMM_UNIFORM, "U(-1,1) operands"))):
    p.gemm(0, warmup=3, reps=10, operands=cls)
    rs = [p.gemm(0, warmup=3, reps=10, operands=cls) for _ in range(N)]
    assert len({(r.c_sum64, r.c_xor32) for r in rs}) == 1, "C checksum changed run to run"
    report(f"GEMM 8192^3 {name} (median of 10)", [r.tflops_median for r in rs], "TFLOP/s")
    report(f"GEMM 8192^3 {name} (best of 10)", [r.tflops_best for r in rs], "TFLOP/s")
if g >= 2:
    S = 256 << 20
    p.nvlink_a2a(list(range(g)), bytes_per_pair=S, warmup=2, reps=5, verify=1)
    rs = [p.nvlink_a2a(list(range(g)), bytes_per_pair=S, warmup=2, reps=10, verify=1) for _ in range(N)]
    report(f"NVLink a2a x{g} 256 MiB egress min over GPUs", [min(r.egress_gbs[:g]) for r in rs], "GB/s")
    report(f"NVLink a2a x{g} 256 MiB pair min", [r.min_pair_gbs for r in rs], "GB/s")
    report(f"NVLink a2a x{g} 256 MiB pair max", [r.max_pair_gbs for r in rs], "GB/s")
    lines.append(f"pair_source of the last call: {rs[-1].pair_source} (0 share, 1 isolated, 2 stepped); pair max/min of the last call {rs[-1].max_pair_gbs / rs[-1].min_pair_gbs:.4f}")
    p.a2a_release()
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/stability_g{g}.txt", "w").write("\n".join(lines) + "\n")
