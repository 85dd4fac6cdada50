"""NVLink a2a tuning on >=2 GPUs: variant x ctas_per_peer x size, single-process probe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
g = torch.cuda.device_count()
ords = list(range(g))
out = []
for mode in (0, 1, 2):
    for S in (1 << 20, 16 << 20, 256 << 20):
        for variant in ((1, 2, 3, 4) if mode != 2 else (0,)):
            try:
                r = p.nvlink_a2a(ords, bytes_per_pair=S, mode=mode, warmup=2, reps=8, variant=variant)
                out.append(f"mode={mode} variant={variant} S={S>>20}MiB ms={r.ms_median:.4f} egress={[round(x,1) for x in r.egress_gbs]} ingress={[round(x,1) for x in r.ingress_gbs]} pair_min={r.min_pair_gbs:.1f} pair_max={r.max_pair_gbs:.1f} verified={r.verified}")
            except Exception as e:
                out.append(f"mode={mode} variant={variant} S={S>>20}MiB ERR {e}")
for variant, cands in ((1, (37, 148)), (2, (18, 37, 74, 148, 296)), (3, (148, 592))):
    for c in cands:
        r = p.nvlink_a2a(ords, bytes_per_pair=256 << 20, mode=0, warmup=2, reps=8, ctas_per_peer=c, variant=variant)
        out.append(f"tune variant={variant} ctas_per_peer={c} egress={[round(x,1) for x in r.egress_gbs]} ingress={[round(x,1) for x in r.ingress_gbs]}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/a2a_tune.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
