"""NVLink all-to-all at the full box (G = all visible GPUs): exchange variants, pair matrix, NCCL leg."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k3s_nvidia_b200.probe import Probe
p = Probe()
g = torch.cuda.device_count()
ords = list(range(g))
rows = {}
S = 256 << 20
for name, kw in (("all_push_tma", dict(mode=0, variant=2)), ("all_pull_tma", dict(mode=0, variant=1)), ("all_push_direct", dict(mode=0, variant=3)),
                 ("nccl_sendrecv", dict(mode=2))):
    try:
        r = p.nvlink_a2a(ords, bytes_per_pair=S, warmup=2, reps=8, **kw)
        rows[name] = dict(ms=round(r.ms_median, 4), egress=[round(x, 1) for x in r.egress_gbs], ingress=[round(x, 1) for x in r.ingress_gbs], verified=r.verified)
    except Exception as e:
        rows[name] = dict(error=str(e)[:200])
    print(name, rows[name], flush=True)
for name, kw in (("pair_pull_64MiB", dict(mode=1, variant=1)), ("pair_push_64MiB", dict(mode=1, variant=2))):
    try:
        r = p.nvlink_a2a(ords, bytes_per_pair=64 << 20, warmup=1, reps=4, **kw)
        rows[name] = dict(min=round(r.min_pair_gbs, 1), max=round(r.max_pair_gbs, 1), matrix=[[round(v) for v in row] for row in r.pair_gbs], verified=r.verified)
    except Exception as e:
        rows[name] = dict(error=str(e)[:200])
    print(name, rows[name], flush=True)
for S2 in (1 << 20, 16 << 20, 1 << 30 if g <= 4 else 512 << 20):
    try:
        r = p.nvlink_a2a(ords, bytes_per_pair=S2, warmup=2, reps=6, mode=0)
        rows[f"all_auto_{S2>>20}MiB"] = dict(ms=round(r.ms_median, 4), egress_min=round(min(r.egress_gbs[:g]), 1), verified=r.verified)
    except Exception as e:
        rows[f"all_auto_{S2>>20}MiB"] = dict(error=str(e)[:200])
    print(S2, rows[f"all_auto_{S2>>20}MiB"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/a2a_g{g}.json", "w"), indent=1)
